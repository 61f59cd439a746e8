"""C3: full learned dynamics -- PlainTransition + PlainRewardMech + PlainTerminationMech, E = 7 / 5 elites,
4 x 200 SiLU, O = 5, A = 1 -- 125 000 fake envs on one GPU (1 M envs over 8 GPUs shard without any collective).
Device-resident steps (CUDA-graph replay, L2 flushed between steps) and host-facing VecFakeEnv.step."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import cmrl_b200 as cm
dev = 'cuda:0'
O, A, N = 5, 1, int(sys.argv[1]) if len(sys.argv) > 1 else 125000
SILU = {"_target_": "torch.nn.SiLU"}
torch.manual_seed(0); rng = np.random.default_rng(0)
tr = cm.PlainTransition(O, A, hid_size=200, activation_fn_cfg=SILU, device=dev)
rw = cm.PlainRewardMech(O, A, hid_size=200, activation_fn_cfg=SILU, device=dev)
tm = cm.PlainTerminationMech(O, A, hid_size=200, activation_fn_cfg=SILU, device=dev)
for m in (tr, rw, tm): m.set_elite_members([0, 1, 2, 3, 4])
with torch.no_grad():  # a termination head that fires rarely (the reference's rule is "non-zero => done": SURVEY section 0)
    tm.mean_and_logvar.weight.zero_(); tm.mean_and_logvar.bias.zero_()
dyn = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw, learned_termination=True, termination_mech=tm)
env = cm.VecFakeEnv(N, None, None)
env.set_up(dyn, get_init_obs_fn=lambda n: rng.normal(size=(n, O)).astype(np.float32), max_episode_steps=1000)
env.rng_mode, env.infos_mode, env.precision, env.deterministic, env.return_views = 'device', 'shared', 'f16', True, True
env.seed(0); env.reset()
act_h = rng.uniform(-1, 1, size=(N, A)).astype(np.float32); act = torch.from_numpy(act_h).to(dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
for _ in range(3): env.step_device_graphed(act)
steps = 50
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
torch.cuda.synchronize()
for a, b in evs:
    flush.fill_(1.0); a.record(); env.step_device_graphed(act); b.record()
torch.cuda.synchronize()
ms = sum(a.elapsed_time(b) for a, b in evs) / steps
print('C3 %d envs, device-resident step: %.3f ms -> %.1f M transitions/s (selected-member FLOPs %.0f per transition -> %.1f TFLOP/s)'
      % (N, ms, N / ms / 1e3, 246400 + 2 * 243200, (246400 + 2 * 243200) * N / ms / 1e9))
for _ in range(3): env.step(act_h)
t0 = time.perf_counter()
for _ in range(steps): env.step(act_h)
dt = (time.perf_counter() - t0) / steps
print('C3 %d envs, VecFakeEnv.step (host numpy in/out, fully learned): %.3f ms -> %.1f M transitions/s' % (N, dt * 1e3, N / dt / 1e6))
