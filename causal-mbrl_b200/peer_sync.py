"""PeerGradSync: the data-parallel gradient exchange of `BaseDynamics.train` over NVLink peer memory.

Host side of `cmrl_allreduce_grads` / `cmrl_allreduce_grads_adam` (csrc/dp_sync.cu): one exchange buffer per rank,
mapped by every peer through CUDA IPC; torch.distributed only carries the 64-byte handles at set-up.  The reference
trains on one device (base_dynamics.py:173-188); this is the `allreduce_grads` entry of SURVEY 8b / the exchange of 8e.
"""
import ctypes

import torch

from . import _lib


class PeerGradSync:
    def __init__(self, group, max_floats: int, device):
        import torch.distributed as dist

        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.max_floats = int(max_floats)
        self.device = torch.device(device)
        self._comm = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(self.device):
            # a rank whose set-up fails at any stage still takes part in both collectives below (with an all-zero handle):
            # leaving early would hang the peers in the all-gather
            err = None
            try:
                _lib.call("cmrl_dp_comm_create", self.rank, self.world, self.max_floats, ctypes.byref(self._comm), handle)
            except _lib.CmrlB200Error as e:  # e.g. no memory for the exchange buffer, CUDA IPC not available in this process
                err = e
            mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(self.device)
            gathered = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(gathered, mine, group=group)
            blob = b"".join(bytes(g.cpu().numpy().tobytes()) for g in gathered)
            if err is None:
                try:
                    _lib.call("cmrl_dp_comm_connect", self._comm, blob)
                except _lib.CmrlB200Error as e:  # e.g. a peer on another node: CUDA IPC only maps memory of the same node
                    err = e
            # every rank learns whether EVERY rank connected (a rank that went on alone would wait for peers that never push);
            # this all-reduce is also the barrier after which pushing into a peer's buffer is safe
            ok = torch.tensor([0 if err is not None else 1], device=self.device, dtype=torch.int32)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            torch.cuda.synchronize(self.device)
            if int(ok.item()) == 0:
                self.close()
                raise _lib.CmrlB200Error("peer-memory gradient exchange: not every rank could map every peer's buffer%s"
                                         % ("" if err is None else " (this rank: %s)" % err))

    def allreduce(self, flat_grad: torch.Tensor):
        """flat_grad <- sum over ranks (in rank order: bit-identical everywhere), one launch on the current stream."""
        assert flat_grad.is_contiguous() and flat_grad.dtype == torch.float32 and flat_grad.numel() <= self.max_floats
        _lib.call("cmrl_allreduce_grads", self._comm, flat_grad.data_ptr(), flat_grad.numel(), torch.cuda.current_stream().cuda_stream)

    def allreduce_adam(self, param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step_dev, grad_scale=1.0):
        """The exchange fused with the L2-Adam update of the whole arena (`step_dev` is incremented)."""
        n = param.numel()
        assert grad.numel() == n and exp_avg.numel() == n and exp_avg_sq.numel() == n and n <= self.max_floats
        _lib.call("cmrl_allreduce_grads_adam", self._comm, param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                  n, lr, beta1, beta2, eps, weight_decay, grad_scale, step_dev.data_ptr(), torch.cuda.current_stream().cuda_stream)

    def check(self):
        """Synchronises; raises if a peer failed to arrive in an exchange (the communicator is then unusable)."""
        err = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            _lib.call("cmrl_dp_comm_error", self._comm, ctypes.byref(err))
        if err.value != 0:
            raise _lib.CmrlB200Error("peer gradient exchange: rank %d did not arrive within 20 s (rank %d waited)"
                                     % (err.value - 1, self.rank))

    def close(self):
        if self._comm:
            _lib.call("cmrl_dp_comm_destroy", self._comm)
            self._comm = ctypes.c_void_p()
