"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): data-parallel training over NCCL equals the single-process
run, replicas stay identical (also on a ragged batch that leaves one rank without rows), sharded hold-out evaluation and
`broadcast_parameters` work over NCCL.  Launches profiles/tools/dp_nccl_check.py under torchrun."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_data_parallel_training_and_evaluation_over_nccl():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29800 + os.getpid() % 100), os.path.join(ROOT, "profiles", "tools", "dp_nccl_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, (res.stdout[-3000:], res.stderr[-3000:])
    out = res.stdout
    assert "weights identical on all ranks: True" in out
    assert "replicas identical" in out and "sharded hold-out evaluation over NCCL" in out
