"""Event trace (clock64 stamps of CTA 0) of the fused training kernels.  Needs the tool build of train_tc.cu:
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --shared -Xcompiler -fPIC -DCMRL_TRACE \
         -o profiles/tools/libcmrl_trace_bin.so causal-mbrl_b200/csrc/api.cu causal-mbrl_b200/csrc/train_tc.cu
Usage: python profiles/tools/train_trace.py [B]"""
import ctypes
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cmrl_b200 as cm  # noqa: E402
from cmrl_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = "cuda:0"
lib = ctypes.CDLL(os.path.join(HERE, "libcmrl_trace_bin.so"))
for name, argtypes in cm._lib.PROTOTYPES.items():
    if name.startswith("cmrl_train"):
        getattr(lib, name).argtypes = argtypes
        getattr(lib, name).restype = ctypes.c_int
lib.cmrl_trace_set.argtypes = [ctypes.c_void_p]
lib.cmrl_last_error.restype = ctypes.c_char_p

model = cm.PlainTransition(5, 1, hid_size=200, activation_fn_cfg={"_target_": "torch.nn.SiLU"}, device=dev)
E = 7
obs = torch.randn(E, B, 5, device=dev)
act = torch.rand(E, B, 1, device=dev)
target = obs + 0.1 * torch.randn(E, B, 5, device=dev)
model.flat_parameters()
plan = ops.fused_train_plan(model, E, B)
mn, mx = model._bounds()
plan.step(obs, act, target, mn, mx, 0.5, 1e-4)  # packs the weights (product library)
torch.cuda.synchronize()
st = ops._stream()
fine = []
mma_fine = []
trace = torch.zeros(5 * 1024, dtype=torch.int64, device=dev)


def run(which, passes=3):
    trace.zero_()
    assert lib.cmrl_trace_set(trace.data_ptr()) == 0
    if which == "fused":
        rc = lib.cmrl_train_fused(plan.packed.data_ptr(), plan.L, plan.P, E, B, 5, 1, plan.hidden, plan.head_dim, plan.layout, plan.act,
                                  plan.residual, passes, obs.data_ptr(), obs.stride(0), act.data_ptr(), act.stride(0), target.data_ptr(),
                                  target.stride(0), 5, mn.data_ptr(), mx.data_ptr(), mn.numel(), 0.5, 1e-4, None, plan.scratch.data_ptr(), st)
    else:
        rc = lib.cmrl_train_dw(plan.scratch.data_ptr(), plan.L, plan.P, E, B, plan.in_size, plan.hidden, plan.head_outputs, plan.act, plan.dw_ptrs,
                               plan.db_ptrs, None, 0, 0, ops._ptr(plan.dw_ws), 0 if plan.dw_ws is None else plan.dw_ws.numel(), st)
    assert rc == 0, lib.cmrl_last_error()
    torch.cuda.synchronize()
    t = trace.cpu().numpy().reshape(5, 1024)
    t0 = min(int(r[r > 0].min()) for r in t if (r > 0).any())
    global fine, mma_fine
    fine = t[3][t[3] > 0] - t0
    mma_fine = t[4][t[4] > 0] - t0
    return [(r[r > 0] - t0) for r in t[:3]]


for passes in (3, 1):
    for _ in range(2):
        prod, epi, mma = run("fused", passes)
    print("=== fused kernel, B=%d, passes=%d (cycles since the first stamp; CTA 0, first tile)" % (B, passes))
    n_ph = 2 * plan.L + 1
    # epilogue warp 0: per phase (wait start, accumulator ready, epilogue done); MMA: per phase (wait start, A ready, first W chunk
    # landed, last W chunk landed, all issued)
    print(" phase | epi: wait-start  acc-ready  done | mma: phase-start  a-first  w-first  w-last  issued")
    for ph in range(min(n_ph, len(epi) // 3)):
        e = epi[3 * ph:3 * ph + 3]
        m = mma[5 * ph:5 * ph + 5] if len(mma) >= 5 * ph + 5 else mma[5 * ph:]
        print("  %2d   | %10d %10d %10d | %s" % (ph, e[0], e[1], e[2], " ".join("%9d" % x for x in m)))
    print(" weight producer: chunk issue times (first 60):", " ".join(str(int(x)) for x in prod[:60]))
    print(" epilogue warp 0, phase 1, per chunk: loop-top | +ldtm | +save z | +activation | +ring acquire | +split store | +publish")
    for i in range(0, len(fine) - 6, 7):
        f = fine[i:i + 7]
        print("    top %7d: %s" % (f[0], " ".join("%5d" % (f[j] - f[j - 1]) for j in range(1, 7))))

    print(" MMA issuer, phase 1, per chunk: loop-top | +A chunk ready | +weights ready | +fences | +MMAs issued | +commits")
    for i in range(0, len(mma_fine) - 5, 6):
        f = mma_fine[i:i + 6]
        print("    top %7d: %s" % (f[0], " ".join("%5d" % (f[j] - f[j - 1]) for j in range(1, 6))))

# attribution experiments (results invalid): which resource paces the fused kernel?
lib.cmrl_trace_flags.argtypes = [ctypes.c_int]
print("=== fused kernel, B=%d, experiments: cycles from the first to the last stamp of CTA 0" % B)
for flags, what in ((0, "everything"), (1, "no weight copies"), (2, "no MMAs"), (4, "no global loads/stores in the epilogue"),
                    (8, "no A-ring writes"), (1 | 4, "no weight copies, no global epilogue traffic"), (2 | 4, "no MMAs, no global epilogue traffic"),
                    (1 | 2, "no weight copies, no MMAs"), (1 | 2 | 4 | 8, "barriers + TMEM loads + activation math only")):
    assert lib.cmrl_trace_flags(flags) == 0
    for passes in (3, 1):
        for _ in range(2):
            prod, epi, mma = run("fused", passes)
        last = max(int(r.max()) for r in (prod, epi, mma) if len(r))
        print("  flags %2d passes %d: %7d cycles   (%s)" % (flags, passes, last, what))
assert lib.cmrl_trace_flags(0) == 0

for _ in range(2):
    prod, tr, mma = run("dw")
print("=== dw kernel, B=%d: CTA (0,0)" % B)
print(" producer issue:", " ".join(str(int(x)) for x in prod[:16]))
print(" transposer warp 0 (per slice: loop-top, raw-ready, stage-free, done) + final accumulator-ready:")
for i in range(0, min(len(tr), 40), 4):
    print("   ", " ".join("%8d" % x for x in tr[i:i + 4]))
print(" mma (operands ready per slice):", " ".join(str(int(x)) for x in mma[:16]))
