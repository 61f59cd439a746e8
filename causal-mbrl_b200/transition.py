"""Transition mechanisms with the reference's constructor signatures, attributes and state_dict layout:
BaseTransition (cmrl/models/transition/base_transition.py:8-26), PlainTransition
(one_step/plain_transition.py:13-122), ExternalMaskTransition (one_step/external_mask_transition.py:15-169)
and ForwardEulerTransition (multi_step/forward_euler.py:8-41).  `forward` runs on libcmrl_b200 kernels."""
from typing import Dict, Optional, Sequence, Tuple, Union

import torch
from torch import nn as nn

from . import ops
from .layers import ParallelEnsembleLinearLayer
from .mlp import EnsembleMLP
from .util import to_tensor


class BaseTransition(EnsembleMLP):
    _MODEL_FILENAME = "base_ensemble_transition.pth"

    def __init__(self, obs_size: int, action_size: int, deterministic: bool, ensemble_num: int = 7,
                 elite_num: int = 5, device: Union[str, torch.device] = "cpu"):  # fmt: skip
        super(BaseTransition, self).__init__(ensemble_num=ensemble_num, elite_num=elite_num, device=device)
        self.obs_size = obs_size
        self.action_size = action_size
        self.deterministic = deterministic

    def forward(self, state: torch.Tensor, action: torch.Tensor):
        pass


class PlainTransition(BaseTransition):
    """Ensemble of MLPs, each a Gaussian over the next observation: 4 x (EnsembleLinear + act) -> head 2*O,
    soft-clamped logvar, residual mean."""

    _MODEL_FILENAME = "plain_transition.pth"

    def __init__(
        self,
        obs_size: int,
        action_size: int,
        deterministic: bool = False,
        ensemble_num: int = 7,
        elite_num: int = 5,
        residual: bool = True,
        learn_logvar_bounds: bool = False,
        num_layers: int = 4,
        hid_size: int = 200,
        activation_fn_cfg: Optional[Dict] = None,
        device: Union[str, torch.device] = "cpu",
    ):
        super().__init__(obs_size=obs_size, action_size=action_size, deterministic=deterministic,
                         ensemble_num=ensemble_num, elite_num=elite_num, device=device)  # fmt: skip
        self.residual = residual
        self.learn_logvar_bounds = learn_logvar_bounds
        self.num_layers = num_layers
        self.hid_size = hid_size
        self.activation_fn_cfg = activation_fn_cfg

        self._build_stack(obs_size + action_size, hid_size, num_layers, obs_size if deterministic else 2 * obs_size,
                          activation_fn_cfg)  # fmt: skip
        if not deterministic:
            self.min_logvar = nn.Parameter(-10 * torch.ones(1, obs_size), requires_grad=learn_logvar_bounds)
            self.max_logvar = nn.Parameter(0.5 * torch.ones(1, obs_size), requires_grad=learn_logvar_bounds)
        self._head_dim = obs_size
        self._finish_init()

    def _residual(self):
        return self.residual

    def forward(
        self,
        batch_obs: torch.Tensor,  # [E, B, obs_size]
        batch_action: torch.Tensor,  # [E, B, action_size]
    ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        assert len(batch_obs.shape) == 3 and batch_obs.shape[-1] == self.obs_size
        assert len(batch_action.shape) == 3 and batch_action.shape[-1] == self.action_size
        return self._mech_forward(batch_obs, batch_action)


class ExternalMaskTransition(BaseTransition):
    """One sub-network per output dimension (P = obs_size parallel nets x E members), each seeing the
    input through its row of a causal-graph mask; head 2 per sub-network.  The repeated / masked input of
    the reference ([P,E,B,in]) is never materialised: the mask is applied in the first GEMM's operand load."""

    _MODEL_FILENAME = "external_mask_transition.pth"
    _head_layout = ops.HEAD_PARALLEL

    def __init__(
        self,
        obs_size: int,
        action_size: int,
        deterministic: bool = False,
        ensemble_num: int = 7,
        elite_num: int = 5,
        residual: bool = True,
        learn_logvar_bounds: bool = False,
        num_layers: int = 4,
        hid_size: int = 200,
        activation_fn_cfg: Optional[Dict] = None,
        device: Union[str, torch.device] = "cpu",
    ):
        super().__init__(obs_size=obs_size, action_size=action_size, deterministic=deterministic,
                         ensemble_num=ensemble_num, elite_num=elite_num, device=device)  # fmt: skip
        self.residual = residual
        self.learn_logvar_bounds = learn_logvar_bounds
        self.num_layers = num_layers
        self.hid_size = hid_size
        self.activation_fn_cfg = activation_fn_cfg

        self._input_mask: Optional[torch.Tensor] = torch.ones((obs_size, obs_size + action_size)).to(device)
        self.add_save_attr("input_mask")

        self._build_stack(obs_size + action_size, hid_size, num_layers, 1 if deterministic else 2, activation_fn_cfg)
        if not deterministic:
            self.min_logvar = nn.Parameter(-10 * torch.ones(obs_size, 1, 1, 1), requires_grad=learn_logvar_bounds)
            self.max_logvar = nn.Parameter(0.5 * torch.ones(obs_size, 1, 1, 1), requires_grad=learn_logvar_bounds)
        self._head_dim = obs_size
        self._finish_init()

    def create_linear_layer(self, l_in, l_out):
        return ParallelEnsembleLinearLayer(l_in, l_out, parallel_num=self.obs_size, ensemble_num=self.ensemble_num)

    def set_input_mask(self, mask):
        self._input_mask = to_tensor(mask).to(self.device)
        self._mask_f32 = None
        self._packed.clear()  # packed W1 caches fold the mask in

    @property
    def input_mask(self):
        return self._input_mask

    def _parallel_num(self):
        return self.obs_size

    def _residual(self):
        return self.residual

    def _kernel_mask(self):
        m = getattr(self, "_mask_f32", None)
        if m is None or m.device != self.mean_and_logvar.weight.device:
            assert self._input_mask is not None
            assert 2 <= self._input_mask.ndim <= 4
            m = self._input_mask.to(device=self.mean_and_logvar.weight.device, dtype=torch.float32).contiguous()
            self._mask_f32 = m
        return m

    def _broadcast_mask(self, E, B):
        m = self._kernel_mask()
        if m.dim() == 2:
            return m[:, None, None, :]
        if m.dim() == 3:
            return m[:, :, None, :] if m.shape[1] == E else m[:, None, :, :]
        return m

    def mask_input(self, x: torch.Tensor) -> torch.Tensor:
        """Reference-compatible helper (external_mask_transition.py:116-139); the forward pass does not call
        it -- the kernels read the mask directly."""
        assert x.ndim == 4
        assert self._input_mask is not None and 2 <= self._input_mask.ndim <= 4
        assert x.shape[0] == self._input_mask.shape[0] and x.shape[-1] == self._input_mask.shape[-1]
        if self._input_mask.ndim == 3 and self._input_mask.shape[1] not in (x.shape[1], x.shape[2]):
            raise RuntimeError("input mask shape %a does not match x shape %a" % (self._input_mask.shape, x.shape))
        x *= self._broadcast_mask(x.shape[1], x.shape[2]).to(x.dtype)
        return x

    def forward(
        self,
        batch_obs: torch.Tensor,  # [E, B, obs_size]
        batch_action: torch.Tensor,  # [E, B, action_size]
    ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        assert len(batch_obs.shape) == 3 and batch_obs.shape[-1] == self.obs_size
        assert len(batch_action.shape) == 3 and batch_action.shape[-1] == self.action_size
        m = self._input_mask
        assert m is not None and 2 <= m.ndim <= 4
        assert m.shape[0] == self.obs_size and m.shape[-1] == self.obs_size + self.action_size
        if m.ndim == 3 and m.shape[1] not in (self.ensemble_num, batch_obs.shape[1]):
            raise RuntimeError("input mask shape %a does not match x shape %a" % (
                tuple(m.shape), (self.obs_size, self.ensemble_num, batch_obs.shape[1], m.shape[-1])))  # fmt: skip
        if m.ndim == 4:
            assert tuple(m.shape[1:3]) == (self.ensemble_num, batch_obs.shape[1])
        return self._mech_forward(batch_obs, batch_action)


class ForwardEulerTransition(BaseTransition):
    """k-fold re-application of a one-step transition (forward_euler.py:8-41)."""

    def __init__(self, one_step_transition: BaseTransition, repeat_times: int = 2):
        super().__init__(
            obs_size=one_step_transition.obs_size,
            action_size=one_step_transition.action_size,
            ensemble_num=one_step_transition.ensemble_num,
            deterministic=one_step_transition.deterministic,
            device=one_step_transition.device,
        )
        self.one_step_transition = one_step_transition
        self.repeat_times = repeat_times

        if hasattr(self.one_step_transition, "max_logvar"):
            self.max_logvar = one_step_transition.max_logvar
            self.min_logvar = one_step_transition.min_logvar
        if hasattr(self.one_step_transition, "input_mask"):
            self.input_mask = self.one_step_transition.input_mask
            self.set_input_mask = self.one_step_transition.set_input_mask

    def set_elite_members(self, elite_indices: Sequence[int]):
        self.one_step_transition.set_elite_members(elite_indices)

    def forward(self, batch_obs: torch.Tensor, batch_action: torch.Tensor):
        logvar = torch.zeros(batch_obs.shape, device=self.device)
        mean = batch_obs
        for _ in range(self.repeat_times):
            mean, logvar = self.one_step_transition.forward(mean, batch_action)
        return mean, logvar

    # ------------------------------------------------------------------ the wrapper owns no layers of its own:
    # parameters, arena, packed copies and kernel-facing attributes are the one-step model's.  Training goes through
    # `get_nll_loss` -> `forward` (autograd over the k applications, BaseDynamics.train_step), rollouts re-apply the
    # one-step kernels k times on the drawn member (VecFakeEnv.predict_selected).
    def flat_parameters(self):
        return self.one_step_transition.flat_parameters()

    def _flatten_parameters(self):
        return self.one_step_transition._flatten_parameters()

    @property
    def _n_trainable(self):
        return self.one_step_transition._n_trainable

    def weights_token(self) -> int:
        return self.one_step_transition.weights_token()

    def _bounds(self):
        return self.one_step_transition._bounds()

    def _kernel_mask(self):
        return self.one_step_transition._kernel_mask()

    def _parallel_num(self):
        return self.one_step_transition._parallel_num()

    def _residual(self):
        return self.one_step_transition._residual()

    @property
    def _head_dim(self):
        return self.one_step_transition._head_dim

    @property
    def _head_layout(self):
        return self.one_step_transition._head_layout

    def packed_f16(self):
        return self.one_step_transition.packed_f16()
