"""EnsembleLinearLayer / ParallelEnsembleLinearLayer with the reference's parameter names, shapes and
initialisation (cmrl/models/layers.py:8-117); `forward` runs the K1 kernel of libcmrl_b200 instead of
`x.matmul(W) + b`."""
import numpy as np
import torch
from torch import nn as nn

from . import ops
from . import util as model_util


class _EnsLinearFn(torch.autograd.Function):
    """y = x @ W + b per net, forward and backward on the K1 kernels."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        y = ops.ens_linear_fwd(x, weight, bias, ops.ACT_IDENTITY)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gy = gy.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = ops.ens_linear_bwd_dx(gy, weight, None, ops.ACT_IDENTITY)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            gw, gb = ops.ens_linear_bwd_dw(x.expand(gy.shape[:-1] + (x.shape[-1],)), gy, weight.shape, want_db=ctx.has_bias)
        return gx, gw, gb


def truncated_normal_init(m: nn.Module):
    """Truncated-normal weights with std 1/(2*sqrt(in)), zero bias, member slice by member slice
    (same RNG order as layers.py:8-28)."""
    if isinstance(m, nn.Linear):
        # the reference reads shape[0] (= out_features) as the "input" dim here (quirk 6, kept)
        stddev = 1 / (2 * np.sqrt(m.weight.data.shape[0]))
        model_util.truncated_normal_(m.weight.data, std=stddev)
        m.bias.data.fill_(0.0)
    elif isinstance(m, EnsembleLinearLayer):
        num_members, input_dim, _ = m.weight.data.shape
        stddev = 1 / (2 * np.sqrt(input_dim))
        for i in range(num_members):
            model_util.truncated_normal_(m.weight.data[i], std=stddev)
        m.bias.data.fill_(0.0)
    elif isinstance(m, ParallelEnsembleLinearLayer):
        num_parallel, num_members, input_dim, _ = m.weight.data.shape
        stddev = 1 / (2 * np.sqrt(input_dim))
        for i in range(num_parallel):
            for j in range(num_members):
                model_util.truncated_normal_(m.weight.data[i, j], std=stddev)
        m.bias.data.fill_(0.0)


class EnsembleLinearLayer(nn.Module):
    """An ensemble of linear layers: weight [E,in,out], bias [E,1,out]."""

    def __init__(self, in_size: int, out_size: int, use_bias: bool = True, ensemble_num: int = 1):
        super().__init__()
        self.ensemble_num = ensemble_num
        self.in_size = in_size
        self.out_size = out_size
        self.weight = nn.Parameter(torch.rand(self.ensemble_num, self.in_size, self.out_size))
        self.use_bias = bool(use_bias)
        if use_bias:
            self.bias = nn.Parameter(torch.rand(self.ensemble_num, 1, self.out_size))

    def forward(self, x):
        return _EnsLinearFn.apply(x, self.weight, self.bias if self.use_bias else None)

    def __repr__(self) -> str:
        return "in_size=%d, out_size=%d, use_bias=%s, ensemble_num=%d" % (
            self.in_size, self.out_size, self.use_bias, self.ensemble_num)  # fmt: skip


class ParallelEnsembleLinearLayer(nn.Module):
    """P parallel sub-networks x E members: weight [P,E,in,out], bias [P,E,1,out]."""

    def __init__(self, in_size: int, out_size: int, use_bias: bool = True, parallel_num: int = 1,
                 ensemble_num: int = 1):  # fmt: skip
        super().__init__()
        self.parallel_num = parallel_num
        self.ensemble_num = ensemble_num
        self.in_size = in_size
        self.out_size = out_size
        self.weight = nn.Parameter(torch.rand(self.parallel_num, self.ensemble_num, self.in_size, self.out_size))
        self.use_bias = bool(use_bias)
        if use_bias:
            self.bias = nn.Parameter(torch.rand(self.parallel_num, self.ensemble_num, 1, self.out_size))

    def forward(self, x):
        return _EnsLinearFn.apply(x, self.weight, self.bias if self.use_bias else None)

    def __repr__(self) -> str:
        return "in_size=%d, out_size=%d, use_bias=%s, parallel_num=%d, ensemble_num=%d" % (
            self.in_size, self.out_size, self.use_bias, self.parallel_num, self.ensemble_num)  # fmt: skip
