"""Conditional-mutual-information test network (cmrl/models/causal_discovery/CMI_test.py:15-130) on the same
grouped masked-GEMM kernels as the transition models.

P = obs_size + action_size + 1 parallel sub-networks x E members; sub-network p sees the input with column p zeroed
(`input_mask = 1 - eye(P, obs_size + action_size)`: the last sub-network sees everything) and predicts ALL obs_size
outputs, so mean / logvar come out as [P, E, B, obs_size] (no transpose, unlike ExternalMaskTransition).  The repeated,
masked input of the reference ([P,E,B,in], CMI_test.py:103-110) is never materialised: the mask is applied in the first
GEMM's operand load.  Parameter names, shapes and the `.pth` file name are the reference's.
"""
from typing import Dict, Optional, Tuple, Union

import torch
from torch import nn as nn

from . import ops
from .layers import ParallelEnsembleLinearLayer
from .mlp import EnsembleMLP, _NllLossFn


class TransitionConditionalMutualInformationTest(EnsembleMLP):
    _MODEL_FILENAME = "conditional_mutual_information_test.pth"
    _head_layout = ops.HEAD_PLAIN

    def __init__(
        self,
        obs_size: int,
        action_size: int,
        ensemble_num: int = 7,
        elite_num: int = 5,
        residual: bool = True,
        learn_logvar_bounds: bool = False,
        num_layers: int = 4,
        hid_size: int = 200,
        activation_fn_cfg: Optional[Dict] = None,
        device: Union[str, torch.device] = "cpu",
    ):
        super().__init__(ensemble_num=ensemble_num, elite_num=elite_num, device=device)
        self.obs_size = obs_size
        self.action_size = action_size
        self.residual = residual
        self.learn_logvar_bounds = learn_logvar_bounds
        self.deterministic = False
        self.num_layers = num_layers
        self.hid_size = hid_size
        self.parallel_num = self.obs_size + self.action_size + 1
        self._input_mask = (1 - torch.eye(self.parallel_num, self.obs_size + self.action_size)).to(self.device)

        self._build_stack(obs_size + action_size, hid_size, num_layers, 2 * obs_size, activation_fn_cfg)
        self.min_logvar = nn.Parameter(-10 * torch.ones(self.parallel_num, 1, 1, obs_size), requires_grad=learn_logvar_bounds)
        self.max_logvar = nn.Parameter(0.5 * torch.ones(self.parallel_num, 1, 1, obs_size), requires_grad=learn_logvar_bounds)
        self._head_dim = obs_size
        self._finish_init()

    def create_linear_layer(self, l_in, l_out):
        return ParallelEnsembleLinearLayer(l_in, l_out, parallel_num=self.parallel_num, ensemble_num=self.ensemble_num)

    @property
    def input_mask(self):
        return self._input_mask

    def mask_input(self, x: torch.Tensor) -> torch.Tensor:
        """Reference-compatible helper (CMI_test.py:89-93); the forward pass reads the mask inside the first GEMM."""
        assert x.ndim == 4
        assert self._input_mask.ndim == 2
        return x * self._input_mask[:, None, None, :]

    # hooks of the shared forward / backward machinery (mlp.py) -------------------------------------------------------
    def _parallel_num(self):
        return self.parallel_num

    def _residual(self):
        return self.residual

    def _kernel_mask(self):
        m = getattr(self, "_mask_f32", None)
        if m is None or m.device != self.mean_and_logvar.weight.device:
            m = self._input_mask.to(device=self.mean_and_logvar.weight.device, dtype=torch.float32).contiguous()
            self._mask_f32 = m
        return m

    def _head_fwd(self, raw, res_obs, E, B):
        """Per sub-net PLAIN head: raw [P,E,B,2O] -> mean, logvar [P,E,B,O]; bounds [P,1,1,O] are per sub-net."""
        P, O = self.parallel_num, self.obs_size
        mean = torch.empty((P, E, B, O), device=raw.device, dtype=torch.float32)
        logvar = torch.empty_like(mean)
        for p in range(P):
            obs_p = None if res_obs is None else (res_obs[p] if res_obs.dim() == 4 else res_obs)
            m_p, lv_p = ops.gaussian_head_fwd(raw[p], ops.HEAD_PLAIN, E, B, O, obs_p, self.min_logvar[p].reshape(O),
                                              self.max_logvar[p].reshape(O))  # fmt: skip
            mean[p], logvar[p] = m_p, lv_p
        return mean, logvar

    def _head_bwd(self, raw, E, B, g_mean, g_logvar, d_min, d_max):
        P, O = self.parallel_num, self.obs_size
        d_raw = torch.empty_like(raw)
        for p in range(P):
            d_raw[p] = ops.gaussian_head_bwd(
                raw[p], ops.HEAD_PLAIN, E, B, O, self.min_logvar[p].reshape(O), self.max_logvar[p].reshape(O), g_mean[p],
                None if g_logvar is None else g_logvar[p], None if d_min is None else d_min[p].reshape(O),
                None if d_max is None else d_max[p].reshape(O))  # fmt: skip
        return d_raw

    def forward(
        self,
        batch_obs: torch.Tensor,  # [(P,) E, B, obs_size]
        batch_action: torch.Tensor,  # [E, B, action_size]
    ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        assert len(batch_action.shape) == 3 and batch_action.shape[-1] == self.action_size
        assert batch_obs.dim() in (3, 4) and batch_obs.shape[-1] == self.obs_size
        if batch_obs.dim() == 4:
            assert batch_obs.shape[0] == self.parallel_num
        return self._mech_forward(batch_obs, batch_action)

    def get_nll_loss(self, model_in: Dict[(str, torch.Tensor)], target: torch.Tensor) -> torch.Tensor:
        """CMI_test.py:123-130: the target is repeated for every sub-net; loss [P,E,B,O]."""
        pred_mean, pred_logvar = self.forward(**model_in)
        P, E, B, O = pred_mean.shape
        target = target.to(pred_mean.device).expand(P, E, B, O).reshape(P * E, B, O)
        learn = self.max_logvar.requires_grad or self.min_logvar.requires_grad
        nll = _NllLossFn.apply(pred_mean.reshape(P * E, B, O), pred_logvar.reshape(P * E, B, O), target,
                               0.0 if learn else self._nll_reg()).view(P, E, B, O)  # fmt: skip
        if learn:
            nll = nll + 0.01 * (self.max_logvar.sum() - self.min_logvar.sum())
        return nll
