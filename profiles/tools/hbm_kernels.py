"""Achieved HBM bandwidth of the memory-bound kernels of the path (K2 head, K3 loss, Adam, K5 tail, draws), each timed
alone inside a CUDA graph (20 launches per replay) at sizes larger than L2, against the measured copy bandwidth in
MEASURED_PEAKS.json.  Algorithmic bytes per unit as in DESIGN.md section 4."""
import json, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cmrl_b200 as cm
dev = 'cuda:0'
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6557.8

def timeit(f, reps=20):
    if os.environ.get("HBM_NCU"):  # under ncu: two eager launches per kernel, the profiler times them
        f(); f(); torch.cuda.synchronize()
        return 1.0
    for _ in range(2): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): f()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3

rows = []
E, B, D = 7, 2_000_000, 5
raw = torch.randn(E, B, 2 * D, device=dev); obs = torch.randn(E, B, D, device=dev)
mn, mx = torch.full((1, D), -10.0, device=dev), torch.full((1, D), 0.5, device=dev)
t = timeit(lambda: cm.ops.gaussian_head_fwd(raw, cm.ops.HEAD_PLAIN, E, B, D, obs, mn, mx))
rows.append(('gaussian_head_fwd', E * B * D * (8 + 4 + 8), t, '[7, 2M, 5]: raw 8 B + obs 4 B read, mean + logvar 8 B written per element'))
mean, logvar = cm.ops.gaussian_head_fwd(raw, cm.ops.HEAD_PLAIN, E, B, D, obs, mn, mx)
tgt = torch.randn(E, B, D, device=dev)
t = timeit(lambda: cm.ops.nll_loss_fwd(mean, logvar, tgt, 0.5))
rows.append(('nll_loss_fwd', E * B * D * 16, t, 'mean, logvar, target read + loss written (16 B per element)'))
t = timeit(lambda: cm.ops.nll_loss_bwd(mean, logvar, tgt, g_scalar=1e-6))
rows.append(('nll_loss_bwd', E * B * D * 20, t, 'mean, logvar, target read + two gradients written (20 B per element)'))
sums = torch.zeros(E, device=dev, dtype=torch.float64)
t = timeit(lambda: cm.ops.mse_member_sums(mean, tgt, sums))
rows.append(('mse_member_sums', E * B * D * 8, t, 'mean, target read (8 B per element), [E] fp64 sums'))
n = 200_000_000
p_, g_, m_, v_ = (torch.randn(n, device=dev) for _ in range(4)); v_.abs_()
t = timeit(lambda: cm.ops.adam_l2_step(p_, g_, m_, v_, 1e-4, 0.9, 0.999, 1e-8, 1e-5, 3), reps=5)
rows.append(('adam_l2_step', n * 28, t, '200 M parameters: p, g, m, v read + p, m, v written (28 B per parameter)'))
del p_, g_, m_, v_
N, O = 20_000_000, 5
nxt, rew = torch.randn(N, O, device=dev), torch.randn(N, 1, device=dev)
term = torch.zeros(N, dtype=torch.uint8, device=dev); env_len = torch.zeros(N, dtype=torch.int32, device=dev)
pool = torch.randn(4096, O, device=dev); ridx = torch.randint(0, 4096, (N,), device=dev, dtype=torch.int32)
o_out, r_out = torch.empty(N, O, device=dev), torch.empty(N, device=dev)
done, trunc, tobs = torch.empty(N, dtype=torch.uint8, device=dev), torch.empty(N, dtype=torch.uint8, device=dev), torch.empty(N, O, device=dev)
t = timeit(lambda: cm.ops.env_step_tail(nxt, rew, None, term, None, 0.0, env_len, 1000, pool, ridx, o_out, r_out, done, trunc, tobs, None), reps=5)
rows.append(('env_step_tail', N * (4 * O + 4 + 1 + 4 + 4 + 4 * O + 4 + 1 + 1 + 4 + 4 * O), t,
             '20 M envs, O=5: next_obs, reward, terminal, env_len, reset index read; obs, reward, done, truncated, env_len, terminal_obs written (83 B per env)'))
elite = torch.tensor([0, 1, 2, 3, 4], dtype=torch.uint8, device=dev)
t = timeit(lambda: cm.ops.draw_member_noise(1, 1, N, elite, O), reps=5)
rows.append(('draw_member_noise', N * (1 + 4 * O), t, '20 M envs: member index 1 B + 5 normals 20 B written per env (Philox + Box-Muller)'))
print('kernel, algorithmic bytes per launch, ms per launch, GB/s, fraction of measured HBM copy bandwidth (%.1f GB/s), shape' % peak)
for name, nbytes, t, note in rows:
    print('%s, %.3e, %.4f, %.1f, %.3f, %s' % (name, nbytes, t * 1e3, nbytes / t / 1e9, nbytes / t / 1e9 / peak, note))
