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
    """Truncated-normal weights with std 1/(2*sqrt(fan)), zero bias; every [in,out] matrix of a grouped layer is drawn
    on its own, net by net in row-major (sub-network, member) order -- the RNG order of layers.py:8-28."""
    if isinstance(m, nn.Linear):
        fan = m.weight.shape[0]  # the reference reads shape[0] (= out_features) as the "input" dim here (quirk 6, kept)
        matrices = [m.weight.data]
    elif isinstance(m, _GroupedLinear):
        fan = m.in_size
        matrices = m.weight.data.reshape((-1,) + tuple(m.weight.shape[-2:]))  # views of the parameter
    else:
        return
    for w in matrices:
        model_util.truncated_normal_(w, std=1 / (2 * np.sqrt(fan)))
    m.bias.data.fill_(0.0)


class _GroupedLinear(nn.Module):
    """`group` = leading dims of the parameters: weight [*group, in, out], bias [*group, 1, out] (uniform(0,1) until
    `truncated_normal_init` is applied, drawn weight first like the reference constructors)."""

    def _build(self, group, in_size: int, out_size: int, use_bias: bool):
        self.in_size = in_size
        self.out_size = out_size
        self.use_bias = bool(use_bias)
        self.weight = nn.Parameter(torch.rand(*group, in_size, out_size))
        if self.use_bias:
            self.bias = nn.Parameter(torch.rand(*group, 1, out_size))

    def forward(self, x):
        return _EnsLinearFn.apply(x, self.weight, self.bias if self.use_bias else None)

    def _describe(self, **group):
        fields = dict(in_size=self.in_size, out_size=self.out_size, use_bias=self.use_bias, **group)
        return ", ".join("%s=%s" % kv for kv in fields.items())


class EnsembleLinearLayer(_GroupedLinear):
    """An ensemble of linear layers: weight [E,in,out], bias [E,1,out] (layers.py:31-68)."""

    def __init__(self, in_size: int, out_size: int, use_bias: bool = True, ensemble_num: int = 1):
        super().__init__()
        self.ensemble_num = ensemble_num
        self._build((ensemble_num,), in_size, out_size, use_bias)

    def __repr__(self) -> str:
        return self._describe(ensemble_num=self.ensemble_num)


class ParallelEnsembleLinearLayer(_GroupedLinear):
    """P parallel sub-networks x E members: weight [P,E,in,out], bias [P,E,1,out] (layers.py:71-117)."""

    def __init__(self, in_size: int, out_size: int, use_bias: bool = True, parallel_num: int = 1,
                 ensemble_num: int = 1):  # fmt: skip
        super().__init__()
        self.parallel_num = parallel_num
        self.ensemble_num = ensemble_num
        self._build((parallel_num, ensemble_num), in_size, out_size, use_bias)

    def __repr__(self) -> str:
        return self._describe(parallel_num=self.parallel_num, ensemble_num=self.ensemble_num)
