import sys, time, json
sys.path.insert(0, '/root/repo')
sys.argv = ['bench.py']
import bench, torch, numpy as np
import cmrl_b200 as cm
ctx = bench._Ctx()
for tier in ("auto", 0):
    cm.ops.TRAIN_GEMM_TIER = tier
    r = bench.time_train_c5(cm, ctx)
    print("tier", tier, r["ms_per_step"], r["tflops_algorithmic"], r["gpu_launches_per_step"])
