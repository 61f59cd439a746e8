"""Per-line wall time of VecFakeEnv.step_wait / _make_infos / step_async (sys.settrace line events, ~1 us overhead per line)."""
import sys, time, linecache, numpy as np, torch
sys.path.insert(0, '/root/repo')
import cmrl_b200 as cm
import cmrl_b200.fake_env as fe
ARGS = sys.argv[1:]
sys.argv = ['bench.py']
import bench
dev = 'cuda:0'
C3 = 'c3' in ARGS  # fully learned step (transition + reward + termination mechs), 125 000 envs
N = 125000 if C3 else 100000
torch.manual_seed(0); np.random.seed(0)
rng = np.random.default_rng(0)
env = cm.VecFakeEnv(N, None, None)
if C3:
    tr = cm.PlainTransition(5, 1, hid_size=200, activation_fn_cfg=bench.SILU, device=dev)
    rw = cm.PlainRewardMech(5, 1, hid_size=200, activation_fn_cfg=bench.SILU, device=dev)
    tm = cm.PlainTerminationMech(5, 1, hid_size=200, activation_fn_cfg=bench.SILU, device=dev)
    with torch.no_grad():
        tm.mean_and_logvar.weight.zero_(); tm.mean_and_logvar.bias.zero_()
    dyn = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw, learned_termination=True, termination_mech=tm)
    env.set_up(dyn, get_init_obs_fn=lambda n: rng.normal(size=(n, 5)).astype(np.float32))
    env.deterministic = True
else:
    tr = cm.ExternalMaskTransition(5, 1, hid_size=200, activation_fn_cfg=bench.SILU, device=dev)
    tr.set_input_mask(torch.from_numpy(bench.make_mask()))
    rw = cm.PlainRewardMech(5, 1, hid_size=200, activation_fn_cfg=bench.SILU, device=dev)
    dyn = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw)
    env.set_up(dyn, termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), dtype=bool), get_init_obs_fn=lambda n: rng.normal(size=(n, 5)).astype(np.float32))
env.rng_mode = 'device'; env.infos_mode = 'shared'; env.precision = 'f16'; env.return_views = True
env.reset()
act = rng.uniform(-1, 1, size=(N, 1)).astype(np.float32)
for _ in range(10): env.step(act)
codes = {fe.VecFakeEnv.step_wait.__code__, fe.VecFakeEnv._make_infos.__code__, fe.VecFakeEnv._predict_all_graphed.__code__,
         fe.VecFakeEnv._predict_all_graphed_one.__code__}
acc = {}
state = {}
def local(frame, event, arg):
    now = time.perf_counter()
    st = state.get(frame)
    if st is not None:
        acc[st[0]] = acc.get(st[0], 0.0) + now - st[1]
    if event == 'return':
        state.pop(frame, None)
        return local
    state[frame] = (frame.f_lineno, time.perf_counter())
    return local
def tracer(frame, event, arg):
    if frame.f_code in codes:
        state[frame] = (frame.f_lineno, time.perf_counter())
        return local
    return None
S = 200
sys.settrace(tracer)
t0 = time.perf_counter()
for _ in range(S): env.step(act)
dt = time.perf_counter() - t0
sys.settrace(None)
print(f"step {dt / S * 1e3:.3f} ms (traced)")
for ln, t in sorted(acc.items(), key=lambda kv: -kv[1])[:22]:
    print(f"{t / S * 1e6:8.1f} us  L{ln}: {linecache.getline(fe.__file__, ln).strip()[:110]}")
