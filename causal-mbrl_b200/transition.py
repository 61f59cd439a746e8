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
        subnet_range: Optional[Tuple[int, int]] = None,
    ):
        """subnet_range=(p0, p1) (not in the reference) builds a NET SHARD: only the sub-networks of output dimensions
        p0..p1-1 exist (parameters [p1-p0, E, ...]); `forward` still takes the full observation and returns
        [E, B, p1-p0].  The sub-networks never exchange anything, so ranks can train disjoint shards with no gradient
        traffic and assemble the full model afterwards with `gather_from_shards`."""
        super().__init__(obs_size=obs_size, action_size=action_size, deterministic=deterministic,
                         ensemble_num=ensemble_num, elite_num=elite_num, device=device)  # fmt: skip
        self.residual = residual
        self.learn_logvar_bounds = learn_logvar_bounds
        self.num_layers = num_layers
        self.hid_size = hid_size
        self.activation_fn_cfg = activation_fn_cfg
        self._sub0, self._sub1 = (0, obs_size) if subnet_range is None else (int(subnet_range[0]), int(subnet_range[1]))
        assert 0 <= self._sub0 < self._sub1 <= obs_size, "subnet_range must be a non-empty range of output dimensions"
        # learned bounds couple the shards through the regulariser 0.01 * (sum(max_logvar) - sum(min_logvar)) over ALL sub-networks
        assert subnet_range is None or not learn_logvar_bounds, "net shards do not support learn_logvar_bounds"
        n_sub = self._sub1 - self._sub0

        self._input_mask: Optional[torch.Tensor] = torch.ones((obs_size, obs_size + action_size)).to(device)
        self.add_save_attr("input_mask")

        self._build_stack(obs_size + action_size, hid_size, num_layers, 1 if deterministic else 2, activation_fn_cfg)
        if not deterministic:
            self.min_logvar = nn.Parameter(-10 * torch.ones(n_sub, 1, 1, 1), requires_grad=learn_logvar_bounds)
            self.max_logvar = nn.Parameter(0.5 * torch.ones(n_sub, 1, 1, 1), requires_grad=learn_logvar_bounds)
        self._head_dim = n_sub
        self._finish_init()

    def create_linear_layer(self, l_in, l_out):
        return ParallelEnsembleLinearLayer(l_in, l_out, parallel_num=self._sub1 - self._sub0, ensemble_num=self.ensemble_num)

    @property
    def is_net_shard(self) -> bool:
        return self._sub1 - self._sub0 != self.obs_size

    @property
    def subnet_range(self) -> Tuple[int, int]:
        return self._sub0, self._sub1

    def _grad_dims(self) -> int:
        """Output dimensions the reference's `loss.mean()` averages over: all of them, also for a shard (so that a shard's
        gradients equal the corresponding slices of the full model's)."""
        return self.obs_size

    def _nll_reg(self) -> float:
        """The loss constant of the FULL model (mlp.py:72-76 sums the bounds of all sub-networks): the bounds are fixed and
        equal across sub-networks (not learned in a shard), so it is this shard's sum scaled up."""
        reg = super()._nll_reg()
        return reg * (self.obs_size / float(self._sub1 - self._sub0)) if self.is_net_shard else reg

    def _head_fwd(self, raw, res_obs, E, B):
        if res_obs is not None and self.is_net_shard:  # residual: this shard's columns of the observation
            res_obs = res_obs[..., self._sub0:self._sub1].contiguous()
        return super()._head_fwd(raw, res_obs, E, B)

    def target_columns(self, target):
        """The columns of a [.., obs_size] training / hold-out target this model predicts."""
        return target[..., self._sub0:self._sub1] if self.is_net_shard else target

    def gather_from_shards(self, group=None) -> "ExternalMaskTransition":
        """Collective over `group`, whose ranks hold consecutive shards covering all output dimensions in rank order:
        returns the full model (every parameter all-gathered along the sub-network axis, mask and elite set of this rank)."""
        import torch.distributed as dist

        world = dist.get_world_size(group)
        full = ExternalMaskTransition(self.obs_size, self.action_size, deterministic=self.deterministic, ensemble_num=self.ensemble_num,
                                      elite_num=self.elite_num, residual=self.residual, learn_logvar_bounds=self.learn_logvar_bounds,
                                      num_layers=self.num_layers, hid_size=self.hid_size, activation_fn_cfg=self.activation_fn_cfg,
                                      device=self.device)  # fmt: skip
        sizes = [torch.zeros(1, dtype=torch.int64, device=self.min_logvar.device if not self.deterministic else self.device) for _ in range(world)]
        mine = torch.tensor([self._sub1 - self._sub0], dtype=torch.int64, device=sizes[0].device)
        dist.all_gather(sizes, mine, group=group)
        sizes = [int(x.item()) for x in sizes]
        assert sum(sizes) == self.obs_size, "the shards of the group do not cover the output dimensions exactly once"
        biggest = max(sizes)
        with torch.no_grad():
            for p_full, p_loc in zip(full.parameters(), self.parameters()):
                pad = torch.zeros((biggest,) + tuple(p_loc.shape[1:]), dtype=p_loc.dtype, device=p_loc.device)
                pad[: p_loc.shape[0]].copy_(p_loc)
                parts = [torch.empty_like(pad) for _ in range(world)]
                dist.all_gather(parts, pad, group=group)
                p_full.copy_(torch.cat([part[:n] for part, n in zip(parts, sizes)], dim=0))
        full.set_input_mask(self._input_mask)
        full.set_elite_members(list(self.elite_members))
        return full

    def set_input_mask(self, mask):
        self._input_mask = to_tensor(mask).to(self.device)
        self._mask_f32 = None
        self._packed.clear()  # packed W1 caches fold the mask in

    @property
    def input_mask(self):
        return self._input_mask

    def _parallel_num(self):
        return self._sub1 - self._sub0  # obs_size unless this is a net shard

    def _residual(self):
        return self.residual

    def _kernel_mask(self):
        m = getattr(self, "_mask_f32", None)
        if m is None or m.device != self.mean_and_logvar.weight.device:
            assert self._input_mask is not None
            assert 2 <= self._input_mask.ndim <= 4
            m = self._input_mask.to(device=self.mean_and_logvar.weight.device, dtype=torch.float32)
            m = m[self._sub0:self._sub1].contiguous()  # (all rows unless this is a net shard)
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
