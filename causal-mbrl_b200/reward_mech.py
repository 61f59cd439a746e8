"""Reward mechanisms: BaseRewardMech (cmrl/models/reward_mech/base_reward_mech.py:8-26) and PlainRewardMech
(plain_reward_mech.py:13-94): same MLP stack as PlainTransition, head 2, scalar logvar bounds, no residual."""
from typing import Dict, Optional, Tuple, Union

import torch
from torch import nn as nn

from .mlp import EnsembleMLP


class BaseRewardMech(EnsembleMLP):
    _MODEL_FILENAME = "base_reward_mech.pth"

    def __init__(self, obs_size: int, action_size: int, deterministic: bool = False, ensemble_num: int = 7,
                 elite_num: int = 5, device: Union[str, torch.device] = "cpu"):  # fmt: skip
        super(BaseRewardMech, self).__init__(ensemble_num=ensemble_num, elite_num=elite_num, device=device)
        self.obs_size = obs_size
        self.action_size = action_size
        self.deterministic = deterministic

    def forward(self, state: torch.Tensor, action: torch.Tensor):
        pass


class _ScalarHeadMixin:
    """Shared body of PlainRewardMech / PlainTerminationMech (byte-identical structure in the reference)."""

    def _init_scalar_mech(self, obs_size, action_size, deterministic, learn_logvar_bounds, num_layers, hid_size,
                          activation_fn_cfg):  # fmt: skip
        self.num_layers = num_layers
        self.hid_size = hid_size
        self._build_stack(obs_size + action_size, hid_size, num_layers, 1 if deterministic else 2, activation_fn_cfg)
        if not deterministic:
            self.min_logvar = nn.Parameter(-10 * torch.ones(1), requires_grad=learn_logvar_bounds)
            self.max_logvar = nn.Parameter(0.5 * torch.ones(1), requires_grad=learn_logvar_bounds)
        self._head_dim = 1
        self._finish_init()

    def _scalar_forward(self, batch_obs, batch_action):
        assert len(batch_obs.shape) == 3 and batch_obs.shape[-1] == self.obs_size
        assert len(batch_action.shape) == 3 and batch_action.shape[-1] == self.action_size
        return self._mech_forward(batch_obs, batch_action)


class PlainRewardMech(_ScalarHeadMixin, BaseRewardMech):
    _MODEL_FILENAME = "plain_reward_mech.pth"

    def __init__(
        self,
        obs_size: int,
        action_size: int,
        deterministic: bool = False,
        ensemble_num: int = 7,
        elite_num: int = 5,
        learn_logvar_bounds: bool = False,
        num_layers: int = 4,
        hid_size: int = 200,
        activation_fn_cfg: Optional[Dict] = None,
        device: Union[str, torch.device] = "cpu",
    ):
        BaseRewardMech.__init__(self, obs_size=obs_size, action_size=action_size, deterministic=deterministic,
                                ensemble_num=ensemble_num, elite_num=elite_num, device=device)  # fmt: skip
        self._init_scalar_mech(obs_size, action_size, deterministic, learn_logvar_bounds, num_layers, hid_size,
                               activation_fn_cfg)  # fmt: skip

    def forward(self, batch_obs: torch.Tensor, batch_action: torch.Tensor) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        return self._scalar_forward(batch_obs, batch_action)
