"""Time the fused rollout kernel of the transition mech (C2, 100k envs) alone for each kernel variant, and print the
role timers of the last one.  Usage: python profiles/tools/variant_time.py 5 4 3"""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, '/root/repo')
import cmrl_b200 as cm
variants = [v for v in sys.argv[1:]] or ['5', '4', '3']  # 'variant[:dbg_flags]'
sys.argv = ['bench.py']
import bench
dev = 'cuda:0'
torch.manual_seed(0); np.random.seed(0)
N = 100000
tr = cm.ExternalMaskTransition(5, 1, hid_size=200, activation_fn_cfg=bench.SILU, device=dev)
tr.set_input_mask(torch.from_numpy(bench.make_mask()))
tr.set_elite_members([0, 1, 2, 3, 4])
dyn = cm.PlainEnsembleDynamics(tr, learned_reward=False)
env = cm.VecFakeEnv(N, None, None)
rng = np.random.default_rng(0)
env.set_up(dyn, reward_fn=lambda n, o, a: np.zeros((n.shape[0], 1), np.float32),
           termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), bool),
           get_init_obs_fn=lambda n: rng.normal(size=(n, 5)).astype(np.float32))
env.precision = 'f16'; env.reset()
act = torch.rand(N, 1, device=dev) * 2 - 1
idx, noise = cm.ops.draw_member_noise(1, 1, N, env._elite_u8('transition'), 5)
env.precision = 'fp32'
ref = env.predict_selected(tr, env._obs_dev, act, idx, noise).clone()
env.precision = 'f16'
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
lib = cm._lib.load(); lib.cmrl_debug_set_buffer.argtypes = [ctypes.c_void_p]
perm, offsets = cm.ops.member_bucket(idx, 7)
for v in variants:
    v, flags = (v.split(':') + ['0'])[:2]
    v, flags = int(v), int(flags)
    lib.cmrl_debug_set_flags(flags)
    cm.ops.TC_VARIANT = v
    tr._packed.clear()
    for _ in range(3): out = env.predict_selected(tr, env._obs_dev, act, idx, noise)
    err = float((out - ref).abs().max() / ref.abs().max())
    pack = tr.packed_f16(); mn, mx = tr._bounds()
    tot = 0.0
    reps = 30
    for r in range(reps):
        flush.fill_(r)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        cm.ops.mlp_rollout_f16(pack, 5, 7, env._obs_dev, act, perm, offsets, 200, 5, cm.ops.HEAD_PARALLEL if hasattr(cm.ops, 'HEAD_PARALLEL') else 1,
                               True, tr._act_kind, True, mn, mx, noise) if False else env.predict_selected(tr, env._obs_dev, act, idx, noise)
        b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    print('variant %d flags %d: predict_selected (bucket + fused kernel) %.4f ms, rel err vs fp32 tier %.2e' % (v, flags, tot / reps, err))
    dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    lib.cmrl_debug_set_buffer(dbg.data_ptr())
    env.predict_selected(tr, env._obs_dev, act, idx, noise); torch.cuda.synchronize()
    lib.cmrl_debug_set_buffer(None); lib.cmrl_debug_set_flags(0)
    d = dbg.cpu().numpy().reshape(148, 16)
    lead = d[0::2]
    items = max(1.0, lead[:, 4].mean())
    print('  MMA slot0 (leader): total, wait_x, wait_a, issue | items', lead[:, :4].mean(0).astype(int), int(items), ' per item:', (lead[:, :4].mean(0) / items).astype(int))
    print('  per-cluster total cycles: min %d mean %d max %d; items min %d max %d' % (lead[:, 0].min(), lead[:, 0].mean(), lead[:, 0].max(), lead[:, 4].min(), lead[:, 4].max()))
    if lead[:, 5].max() > 0: print('  SM clock during the kernel: %.0f MHz (cycles / globaltimer ns)' % (1e3 * lead[:, 0].sum() / lead[:, 5].sum()))
    if d[:, 6].max() > 0:
        st, en = d[:, 6], d[:, 7]
        print('  wall clock (globaltimer): first CTA entry -> last CTA exit %.1f us; entry spread %.1f us; exit spread %.1f us; mean CTA lifetime %.1f us'
              % ((en.max() - st.min()) / 1e3, (st.max() - st.min()) / 1e3, (en.max() - en.min()) / 1e3, (en - st).mean() / 1e3))
    print('  epilogue warp0: total, wait_d, loop, -, -, -, publish', d[:, 8:15].mean(0).astype(int), ' per item:', (d[:, 8:15].mean(0) / items).astype(int))
