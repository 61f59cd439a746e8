"""Small wide model for an ncu capture of the streamed-weights kernel (ExternalMask O = P = 8, A = 4, E = 7, H = 512, 40 000 envs):
the full C5 model holds 11 GB of parameters and arenas, which ncu would save / restore around every replay pass."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import cmrl_b200 as cm
dev = 'cuda:0'
O, A, E, H, N = 8, 4, 7, 512, 40000
torch.manual_seed(0); rng = np.random.default_rng(0)
m = cm.ExternalMaskTransition(O, A, ensemble_num=E, hid_size=H, activation_fn_cfg={"_target_": "torch.nn.SiLU"}, device=dev)
m.set_elite_members([0, 1, 2, 3, 4])
dyn = cm.PlainEnsembleDynamics(m, learned_reward=False)
env = cm.VecFakeEnv(N, None, None)
env.set_up(dyn, reward_fn=lambda n, o, a: np.zeros((n.shape[0], 1), np.float32), termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), bool),
           get_init_obs_fn=lambda n: rng.normal(size=(n, O)).astype(np.float32))
env.reset(); env.precision = 'f16'
act = torch.rand(N, A, device=dev) * 2 - 1
idx, noise = cm.ops.draw_member_noise(1, 1, N, env._elite_u8('transition'), O)
for _ in range(3): env.predict_selected(m, env._obs_dev, act, idx, noise)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); env.predict_selected(m, env._obs_dev, act, idx, noise); b.record(); torch.cuda.synchronize()
flop = 2.0 * (12 * 512 + 3 * 512 * 512 + 512 * 2) * O * N
print('v6: %.3f ms, %.1f TFLOP/s algorithmic' % (a.elapsed_time(b), flop / a.elapsed_time(b) / 1e9))
