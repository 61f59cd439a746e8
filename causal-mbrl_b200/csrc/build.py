"""Build libcmrl_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python causal-mbrl_b200/csrc/build.py [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libcmrl_b200.so")
STAMP = os.path.join(PKG, ".libcmrl_b200.stamp")
SOURCES = ["api.cu", "ens_linear.cu", "head_loss.cu", "rollout.cu", "mlp_tc.cu", "mlp_tc3.cu", "mlp_tc5.cu", "mlp_tc6.cu", "gemm_tc.cu", "train_tc.cu", "dp_sync.cu"]
NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-lineinfo",
    "-std=c++17",
    "--shared",
    "-Xcompiler",
    "-fPIC",
    "-Xptxas",
    "-v",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _digest(paths):
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    deps = srcs + [os.path.join(HERE, f) for f in sorted(os.listdir(HERE)) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(PKG), "include", "cmrl_b200.h"))
    digest = _digest(deps)
    if not force and os.path.exists(OUT) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return OUT
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", OUT] + srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(PKG, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libcmrl_b200.so")
    if verbose:
        print(log)
    with open(STAMP, "w") as f:
        f.write(digest)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
