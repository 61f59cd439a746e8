"""Fused rollout kernel of the REWARD mech alone (1 net, 100k envs: a short launch) for each kernel variant.
Timed alone (cold L2) the single-CTA kernel v4 wins below ~1 100 tiles (65.8 vs 72.0 us at 100k envs), but choosing it for
short launches made the graph-replayed step SLOWER (C2 0.250 -> 0.263 ms, C3 0.216 -> 0.231 ms): not kept."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import cmrl_b200 as cm
dev = 'cuda:0'
SILU = {"_target_": "torch.nn.SiLU"}
torch.manual_seed(0); np.random.seed(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rw = cm.PlainRewardMech(5, 1, hid_size=200, activation_fn_cfg=SILU, device=dev)
rw.set_elite_members([0, 1, 2, 3, 4])
tr = cm.PlainTransition(5, 1, hid_size=200, activation_fn_cfg=SILU, device=dev)
dyn = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw)
env = cm.VecFakeEnv(N, None, None)
rng = np.random.default_rng(0)
env.set_up(dyn, termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), bool), get_init_obs_fn=lambda n: rng.normal(size=(n, 5)).astype(np.float32))
env.precision = 'f16'; env.reset()
act = torch.rand(N, 1, device=dev) * 2 - 1
idx, noise = cm.ops.draw_member_noise(1, 1, N, env._elite_u8('reward_mech'), 1)
bucket = cm.ops.member_bucket(idx, 7)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
env.precision = 'fp32'
ref = env.predict_selected(rw, env._obs_dev, act, idx, noise, bucket).clone()
env.precision = 'f16'
for v in (5, 4, 3, 1):
    cm.ops.TC_VARIANT = v
    rw._packed.clear()
    for _ in range(3): out = env.predict_selected(rw, env._obs_dev, act, idx, noise, bucket)
    err = float((out - ref).abs().max() / ref.abs().max())
    tot = 0.0
    reps = 30
    for r in range(reps):
        flush.fill_(r)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.predict_selected(rw, env._obs_dev, act, idx, noise, bucket); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    print('variant %d: reward mech kernel %.4f ms (N=%d), rel err %.2e' % (v, tot / reps, N, err))
