"""Host-side helpers mirroring cmrl/models/util.py (gaussian_nll :14-37, truncated_normal_ :42-61,
to_tensor :64-69).  gaussian_nll on CUDA tensors dispatches to the K3 kernel."""
import numpy as np
import torch

from . import ops


def gaussian_nll(pred_mean, pred_logvar, target, reduce=True):
    """(mu - y)^2 * exp(-logvar) + logvar; `reduce=True` -> losses.sum(dim=1).mean() like the reference."""
    from .mlp import _NllLossFn

    losses = _NllLossFn.apply(pred_mean, pred_logvar, target, 0.0)
    if reduce:
        return losses.sum(dim=1).mean()
    return losses


def truncated_normal_(tensor, mean=0, std=1):
    """In-place N(mean, std^2) truncated to +-2 std by resampling the violators until none is left.
    Same RNG call sequence as the reference (normal_ on the slice, then torch.normal(size=(n,)) per
    round), so a given torch seed yields the reference's initial weights."""
    torch.nn.init.normal_(tensor, mean=mean, std=std)
    while True:
        outside = torch.logical_or(tensor < mean - 2 * std, tensor > mean + 2 * std)
        n_bad = int(outside.sum().item())
        if n_bad == 0:
            return tensor
        tensor[outside] = torch.normal(mean, std, size=(n_bad,), device=tensor.device)


def to_tensor(x):
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x)
    raise ValueError("Input must be torch.Tensor or np.ndarray.")
