"""Time the K1 training GEMMs (fwd with pre-activation, dx, dw) per tier: 0 = fp32 FFMA, 1 = TF32, 3 = 3xTF32 (tcgen05)."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import cmrl_b200 as cm
dev = 'cuda:0'
def timeit(f, reps=20):
    """kernel time without the host-side call overhead: `reps` calls captured in one CUDA graph, replayed"""
    for _ in range(3): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): f()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for (E, B, i, o) in [(7, 256, 200, 200), (7, 256, 6, 200), (7, 256, 200, 10), (7, 4096, 200, 200), (7, 65536, 200, 200)]:
    x, w, b, dz = torch.randn(E, B, i, device=dev), torch.randn(E, i, o, device=dev), torch.randn(E, 1, o, device=dev), torch.randn(E, B, o, device=dev)
    z = torch.randn(E, B, i, device=dev)
    row = []
    for tier in (0, 1, 3):
        t_f = timeit(lambda: cm.ops.ens_linear_fwd(x, w, b, cm.ops.ACT_SILU, want_z=True, tier=tier))
        t_x = timeit(lambda: cm.ops.ens_linear_bwd_dx(dz, w, z, cm.ops.ACT_SILU, tier=tier))
        t_w = timeit(lambda: cm.ops.ens_linear_bwd_dw(x, dz, w.shape, tier=tier))
        row.append((tier, t_f, t_x, t_w))
    fl = 2.0 * E * B * i * o
    print('E=%d B=%d in=%d out=%d (%.1f MFLOP):' % (E, B, i, o, fl / 1e6), '  '.join('tier %d fwd %.1f dx %.1f dw %.1f us' % r for r in row),
          '| tier3 fwd %.1f TFLOP/s' % (fl / row[2][1] / 1e6))
