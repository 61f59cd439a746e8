"""C5 (O = P = 100, A = 20, E = 16 / 12 elites, H = 512, per-member mask): selected-member rollout prediction of the
transition mech, fp32 tier (ragged FFMA GEMMs) vs the streamed-weights tcgen05 kernel, N envs."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import cmrl_b200 as cm
dev = 'cuda:0'
O, A, E, H = 100, 20, 16, 512
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
torch.manual_seed(0); rng = np.random.default_rng(0)
m = cm.ExternalMaskTransition(O, A, ensemble_num=E, elite_num=12, hid_size=H, activation_fn_cfg={"_target_": "torch.nn.SiLU"}, device=dev)
mask = (rng.uniform(size=(O, E, O + A)) < 0.3).astype(np.float32); mask[np.arange(O), :, np.arange(O)] = 1.0
m.set_input_mask(torch.from_numpy(mask)); m.set_elite_members(list(range(12)))
dyn = cm.PlainEnsembleDynamics(m, learned_reward=False)
env = cm.VecFakeEnv(N, None, None)
env.set_up(dyn, reward_fn=lambda n, o, a: np.zeros((n.shape[0], 1), np.float32), termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), bool),
           get_init_obs_fn=lambda n: rng.normal(size=(n, O)).astype(np.float32))
env.reset()
act = torch.rand(N, A, device=dev) * 2 - 1
idx, noise = cm.ops.draw_member_noise(1, 1, N, env._elite_u8('transition'), O)
flop = 2.0 * (120 * 512 + 3 * 512 * 512 + 512 * 2) * O * N
res = {}
for prec in ('fp32', 'f16'):
    env.precision = prec
    for _ in range(2): out = env.predict_selected(m, env._obs_dev, act, idx, noise)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): out = env.predict_selected(m, env._obs_dev, act, idx, noise)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    res[prec] = out.clone()
    print('C5 N=%d %s: %.2f ms per prediction, %.0f transitions/s, %.1f TFLOP/s algorithmic (selected member)' % (N, prec, ms, N / ms * 1e3, flop / ms / 1e9))
print('rel err f16 vs fp32: %.2e' % float((res['f16'] - res['fp32']).abs().max() / res['fp32'].abs().max()))
