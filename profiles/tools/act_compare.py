"""The fused rollout kernel (C2 transition mech, 100k envs) with SiLU (the dynamics yaml), ReLU (the constructor default)
and no activation: how much of the kernel time is the SFU-bound SiLU epilogue."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import cmrl_b200 as cm
sys.argv = ['bench.py']
import bench
dev = 'cuda:0'
N = 100000
flush = torch.empty(64 * 1024 * 1024, device=dev)
for name, cfg in (('SiLU', bench.SILU), ('ReLU', None), ('Identity', {"_target_": "torch.nn.Identity"})):
    torch.manual_seed(0); rng = np.random.default_rng(0)
    tr = cm.ExternalMaskTransition(5, 1, hid_size=200, activation_fn_cfg=cfg, device=dev)
    tr.set_input_mask(torch.from_numpy(bench.make_mask())); tr.set_elite_members([0, 1, 2, 3, 4])
    dyn = cm.PlainEnsembleDynamics(tr, learned_reward=False)
    env = cm.VecFakeEnv(N, None, None)
    env.set_up(dyn, reward_fn=lambda n, o, a: np.zeros((n.shape[0], 1), np.float32), termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), bool),
               get_init_obs_fn=lambda n: rng.normal(size=(n, 5)).astype(np.float32))
    env.precision = 'f16'; env.reset()
    act = torch.rand(N, 1, device=dev) * 2 - 1
    idx, noise = cm.ops.draw_member_noise(1, 1, N, env._elite_u8('transition'), 5)
    perm, offsets = cm.ops.member_bucket(idx, 7)
    mn, mx = tr._bounds(); pack = tr.packed_f16()
    f = lambda: cm.ops.mlp_rollout_f16(pack, 5, 7, env._obs_dev, act, perm, offsets, 200, 5, cm.ops.HEAD_PARALLEL, True, tr._act_kind, True, mn, mx, noise)
    for _ in range(3): f()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    torch.cuda.synchronize()
    for a, b in evs:
        flush.fill_(1.0); a.record(); f(); b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / 20
    print('%-8s: %.4f ms per launch, %.0f TFLOP/s algorithmic, %.3f of the measured bf16 peak (1661.8)' % (name, ms, 121.6 / ms, 121.6 / ms / 1661.8))
