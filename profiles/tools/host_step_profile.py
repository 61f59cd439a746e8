"""Where does the host time of VecFakeEnv.step (e2e path) go?  Wall-clock breakdown with explicit syncs."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import cmrl_b200 as cm
sys.argv = ['bench.py']
import bench
dev = 'cuda:0'
N = 100000
torch.manual_seed(0); np.random.seed(0)
tr = cm.ExternalMaskTransition(5, 1, hid_size=200, activation_fn_cfg=bench.SILU, device=dev)
tr.set_input_mask(torch.from_numpy(bench.make_mask()))
rw = cm.PlainRewardMech(5, 1, hid_size=200, activation_fn_cfg=bench.SILU, device=dev)
dyn = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw)
rng = np.random.default_rng(0)
env = cm.VecFakeEnv(N, None, None)
env.set_up(dyn, termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), dtype=bool), get_init_obs_fn=lambda n: rng.normal(size=(n, 5)).astype(np.float32))
env.rng_mode = 'device'; env.infos_mode = 'shared'; env.precision = 'f16'
env.reset()
act = rng.uniform(-1, 1, size=(N, 1)).astype(np.float32)
for _ in range(5): env.step(act)
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
t0 = time.perf_counter()
for _ in range(50): env.step(act)
dt = time.perf_counter() - t0
pr.disable()
print('ms per step', dt / 50 * 1e3)
pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
