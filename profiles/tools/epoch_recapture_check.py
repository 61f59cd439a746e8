"""Device-resident epochs over one data set replay ONE captured step: the capture count stays at 1 over several epochs and
the per-epoch time is flat (a key of the runner that changed from epoch to epoch would re-capture every time).
    python profiles/tools/epoch_recapture_check.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cmrl_b200 as cm  # noqa: E402
from cmrl_b200 import ops  # noqa: E402

captures = [0]
real_capture = ops.graph_capture


def counting_capture(g):
    captures[0] += 1
    return real_capture(g)


ops.graph_capture = counting_capture
import importlib  # noqa: E402

dyn_mod = importlib.import_module(cm.PlainEnsembleDynamics.__module__)
dyn_mod.ops.graph_capture = counting_capture


class Buf:
    pass


rng = np.random.default_rng(0)
for rows, bs in ((80_000, 256), (123_000, 4096)):
    captures[0] = 0
    b = Buf()
    b.buffer_size, b.full, b.pos = rows, True, 0
    obs = rng.normal(size=(rows, 5)).astype(np.float32)
    b.observations = obs[:, None]
    b.next_observations = (obs + 0.1 * rng.normal(size=(rows, 5)).astype(np.float32))[:, None]
    b.actions = rng.uniform(-1, 1, size=(rows, 1, 1)).astype(np.float32)
    b.rewards = rng.normal(size=(rows, 1)).astype(np.float32)
    b.dones = np.zeros((rows, 1), bool)
    tr = cm.PlainTransition(5, 1, hid_size=200, activation_fn_cfg={"_target_": "torch.nn.SiLU"}, device="cuda:0")
    dyn = cm.PlainEnsembleDynamics(tr, learned_reward=False)
    train_it, _ = dyn.dataset_split(b, 0.2, bs)
    times = []
    for epoch in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss = dyn.train_epoch_device(train_it, "transition")
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    steps = -(-train_it.num_stored // bs)
    print("bs %d: %d steps per epoch, captures %d, epoch times (ms) %s, ms per step (last) %.4f, loss %.4f"
          % (bs, steps, captures[0], ["%.1f" % (1e3 * t) for t in times], 1e3 * times[-1] / steps, loss))
    assert captures[0] == 1, "the step was captured more than once"
