"""Kernel launches for the round-2 ncu captures, eager (no CUDA graphs), three iterations each (ncu: skip the first):
fused training step (pack, train_fused_kernel, train_dw_kernel) at 256 and 4 096 rows per member on the C1/C4 model, the
fp32-grade rollout (train_fused_kernel in rollout mode) and the f16 rollout (mlp_rollout_f16_v5_kernel) of the C2 transition
mech at 100 000 envs, and one device-resident step (draws, buckets, step tail).
    ncu --set full --clock-control none --import-source on -k regex:'train_|mlp_rollout|env_step_tail' -o gpurun_out/r2_full python profiles/tools/r2_ncu_case.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cmrl_b200 as cm  # noqa: E402
from cmrl_b200 import ops  # noqa: E402

dev = "cuda:0"
SILU = {"_target_": "torch.nn.SiLU"}
E, O, A = 7, 5, 1
torch.manual_seed(0)
rng = np.random.default_rng(0)

tr = cm.PlainTransition(O, A, hid_size=200, activation_fn_cfg=SILU, device=dev)
tr.flat_parameters()
mn, mx = tr._bounds()
for B in (256, 4096):
    obs = torch.randn(E, B, O, device=dev)
    act = torch.rand(E, B, A, device=dev) * 2 - 1
    tgt = obs + 0.1 * torch.tanh(obs)
    plan = ops.fused_train_plan(tr, E, B)
    for _ in range(3):
        plan.step(obs, act, tgt, mn, mx, tr._nll_reg(), 1.0 / (E * B * O))
    torch.cuda.synchronize()

N = 100_000
em = cm.ExternalMaskTransition(O, A, hid_size=200, activation_fn_cfg=SILU, device=dev)
mask = (rng.uniform(size=(O, O + A)) < 0.5).astype(np.float32)
mask[np.arange(O), np.arange(O)] = 1
em.set_input_mask(torch.from_numpy(mask))
rw = cm.PlainRewardMech(O, A, hid_size=200, activation_fn_cfg=SILU, device=dev)
em.set_elite_members(list(range(5)))
rw.set_elite_members(list(range(5)))
dyn = cm.PlainEnsembleDynamics(em, learned_reward=True, reward_mech=rw)
env = cm.VecFakeEnv(N, None, None)
env.set_up(dyn, termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), bool), get_init_obs_fn=lambda n: rng.normal(size=(n, O)).astype(np.float32))
env.rng_mode = "device"
env.reset()
obs = env._obs_dev.clone()
act = torch.rand(N, A, device=dev) * 2 - 1
idx = torch.from_numpy(rng.choice(5, size=N).astype(np.uint8)).to(dev)
noise = torch.randn(N, O, device=dev)
for prec in ("fp32", "f16"):
    env.precision = prec
    for _ in range(3):
        env.predict_selected(em, obs, act, idx, noise)
    torch.cuda.synchronize()
env.precision = "f16"
for _ in range(3):
    env.step_device(act)
torch.cuda.synchronize()
