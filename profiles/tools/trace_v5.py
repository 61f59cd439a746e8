"""Event trace of CTA 0 of the v5 kernel: epilogue warp 0 and the two MMA issuers (clock64 stamps)."""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, '/root/repo')
import cmrl_b200 as cm
ARGS = sys.argv[:]
sys.argv = ['bench.py']
import bench
dev = 'cuda:0'
torch.manual_seed(0); np.random.seed(0)
N = 100000
tr = cm.ExternalMaskTransition(5, 1, hid_size=200, activation_fn_cfg=bench.SILU, device=dev)
tr.set_input_mask(torch.from_numpy(bench.make_mask())); tr.set_elite_members([0, 1, 2, 3, 4])
dyn = cm.PlainEnsembleDynamics(tr, learned_reward=False)
env = cm.VecFakeEnv(N, None, None)
rng = np.random.default_rng(0)
env.set_up(dyn, reward_fn=lambda n, o, a: np.zeros((n.shape[0], 1), np.float32), termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), bool), get_init_obs_fn=lambda n: rng.normal(size=(n, 5)).astype(np.float32))
env.precision = 'f16'; env.reset()
act = torch.rand(N, 1, device=dev) * 2 - 1
idx, noise = cm.ops.draw_member_noise(1, 1, N, env._elite_u8('transition'), 5)
cm.ops.TC_VARIANT = 5
for _ in range(3): env.predict_selected(tr, env._obs_dev, act, idx, noise)
lib = cm._lib.load(); lib.cmrl_debug_set_buffer.argtypes = [ctypes.c_void_p]
dbg = torch.zeros(4096 + 3 * 512, dtype=torch.int64, device=dev)
lib.cmrl_debug_set_flags(256 | int(ARGS[1]) if len(ARGS) > 1 else 256); lib.cmrl_debug_set_buffer(dbg.data_ptr())
env.predict_selected(tr, env._obs_dev, act, idx, noise); torch.cuda.synchronize()
lib.cmrl_debug_set_buffer(None); lib.cmrl_debug_set_flags(0)
t = dbg.cpu().numpy()[4096:].reshape(3, 512)
t0 = t[t > 0].min()
e0 = t[0][t[0] > 0] - t0; epi = e0[: (len(e0) // 3) * 3].reshape(-1, 3)[:40]
print('epilogue warp0 tile-steps (start, acc-ready, published; durations: wait, work):')
for i, (a, b, c) in enumerate(epi[16:40]):
    print(f'  step {i+16:3d} slot {i%2}: start {a:7d} ready {b:7d} pub {c:7d} | wait {b-a:5d} work {c-b:5d}')
for r in (1, 2):
    m = (t[r][t[r] > 0] - t0)
    m = m[: (len(m) // 3) * 3].reshape(-1, 3)[8:20]
    print('MMA issuer slot', r - 1, '(wait start, issue start, issue end):')
    for a, b, c in m:
        print(f'  wait {a:7d} -> {b:7d} ({b-a:5d})  issue -> {c:7d} ({c-b:5d})')
