"""Termination mechanisms: BaseTerminationMech (cmrl/models/termination_mech/base_termination_mech.py:8-26,
whose _MODEL_FILENAME is the reference's copy-pasted "base_reward_mech.pth", kept) and PlainTerminationMech
(plain_termination_mech.py:13-94)."""
from typing import Dict, Optional, Tuple, Union

import torch

from .mlp import EnsembleMLP
from .reward_mech import _ScalarHeadMixin


class BaseTerminationMech(EnsembleMLP):
    _MODEL_FILENAME = "base_reward_mech.pth"

    def __init__(self, obs_size: int, action_size: int, deterministic: bool = False, ensemble_num: int = 7,
                 elite_num: int = 5, device: Union[str, torch.device] = "cpu"):  # fmt: skip
        super(BaseTerminationMech, self).__init__(ensemble_num=ensemble_num, elite_num=elite_num, device=device)
        self.obs_size = obs_size
        self.action_size = action_size
        self.deterministic = deterministic

    def forward(self, state: torch.Tensor, action: torch.Tensor):
        pass


class PlainTerminationMech(_ScalarHeadMixin, BaseTerminationMech):
    _MODEL_FILENAME = "plain_termination_mech.pth"

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
        BaseTerminationMech.__init__(self, obs_size=obs_size, action_size=action_size, deterministic=deterministic,
                                     ensemble_num=ensemble_num, elite_num=elite_num, device=device)  # fmt: skip
        self._init_scalar_mech(obs_size, action_size, deterministic, learn_logvar_bounds, num_layers, hid_size,
                               activation_fn_cfg)  # fmt: skip

    def forward(self, batch_obs: torch.Tensor, batch_action: torch.Tensor) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        return self._scalar_forward(batch_obs, batch_action)
