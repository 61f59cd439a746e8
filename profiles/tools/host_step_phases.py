"""Wall-clock phases of VecFakeEnv.step (e2e path): graph launches vs stream waits vs the rest (no profiler)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import cmrl_b200 as cm
sys.argv = ['bench.py']
import bench
dev = 'cuda:0'
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
torch.manual_seed(0); np.random.seed(0)
tr = cm.ExternalMaskTransition(5, 1, hid_size=200, activation_fn_cfg=bench.SILU, device=dev)
tr.set_input_mask(torch.from_numpy(bench.make_mask()))
rw = cm.PlainRewardMech(5, 1, hid_size=200, activation_fn_cfg=bench.SILU, device=dev)
dyn = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw)
rng = np.random.default_rng(0)
env = cm.VecFakeEnv(N, None, None)
env.set_up(dyn, termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), dtype=bool), get_init_obs_fn=lambda n: rng.normal(size=(n, 5)).astype(np.float32))
env.rng_mode = 'device'; env.infos_mode = 'shared'; env.precision = 'f16'; env.return_views = True
env.reset()
act = rng.uniform(-1, 1, size=(N, 1)).astype(np.float32)
for _ in range(10): env.step(act)
acc = {}
def wrap(cls, name, key):
    orig = getattr(cls, name)
    cnt = [0]
    def f(self, *a, **k):
        t0 = time.perf_counter()
        r = orig(self, *a, **k)
        k2 = f"{key}{cnt[0] % 2}"; cnt[0] += 1
        acc[k2] = acc.get(k2, 0.0) + time.perf_counter() - t0
        return r
    setattr(cls, name, f)
wrap(torch.cuda.CUDAGraph, 'replay', 'replay')
wrap(torch.cuda.Stream, 'synchronize', 'sync')
wrap(torch.cuda.Event, 'synchronize', 'evsync')
S = 200
t0 = time.perf_counter()
for _ in range(S): env.step(act)
dt = time.perf_counter() - t0
print(f"step {dt / S * 1e3:.3f} ms:", {k: round(v / S * 1e3, 3) for k, v in sorted(acc.items())}, "rest", round((dt - sum(acc.values())) / S * 1e3, 3))
