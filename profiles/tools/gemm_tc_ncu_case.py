"""One forward / dx / split-K dw call of the tcgen05 3xTF32 training GEMMs at 4096 rows per member (E = 7, 200 x 200) for ncu."""
import sys, torch
sys.path.insert(0, '/root/repo')
import cmrl_b200 as cm
dev = 'cuda:0'
E, B, i, o = 7, 4096, 200, 200
x, w, b, dz, z = (torch.randn(E, B, i, device=dev), torch.randn(E, i, o, device=dev), torch.randn(E, 1, o, device=dev),
                  torch.randn(E, B, o, device=dev), torch.randn(E, B, i, device=dev))
for _ in range(3):
    cm.ops.ens_linear_fwd(x, w, b, cm.ops.ACT_SILU, want_z=True, tier=3)
    cm.ops.ens_linear_bwd_dx(dz, w, z, cm.ops.ACT_SILU, tier=3)
    cm.ops.ens_linear_bwd_dw(x, dz, w.shape, tier=3)
torch.cuda.synchronize()
