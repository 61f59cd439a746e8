"""Batch container mirroring cmrl/types.py:20-77 (InteractionBatch) and its type aliases."""
from dataclasses import dataclass
from typing import Callable, Optional, Tuple, Union

import numpy as np
import torch

RewardFnType = Callable[[torch.Tensor, torch.Tensor, torch.Tensor], torch.Tensor]
TermFnType = Callable[[torch.Tensor, torch.Tensor, torch.Tensor], torch.Tensor]
InitObsFnType = Callable[[int], torch.Tensor]
ObsProcessFnType = Callable[[np.ndarray], np.ndarray]
TensorType = Union[torch.Tensor, np.ndarray]
TrajectoryEvalFnType = Callable[[TensorType, torch.Tensor], torch.Tensor]
InteractionData = Tuple[TensorType, TensorType, TensorType, TensorType, TensorType]

_FIELDS = ("batch_obs", "batch_action", "batch_next_obs", "batch_reward", "batch_done")


@dataclass
class InteractionBatch:
    """A batch of (obs, action, next_obs, reward, done) transitions."""

    batch_obs: Optional[TensorType]
    batch_action: Optional[TensorType]
    batch_next_obs: Optional[TensorType]
    batch_reward: Optional[TensorType]
    batch_done: Optional[TensorType]

    @property
    def attrs(self):
        return list(_FIELDS)

    def __len__(self):
        return self.batch_obs.shape[0]

    def as_tuple(self) -> InteractionData:
        return tuple(getattr(self, f) for f in _FIELDS)

    def __getitem__(self, item):
        return InteractionBatch(*(getattr(self, f)[item] for f in _FIELDS))

    @staticmethod
    def _get_new_shape(old_shape: Tuple[int, ...], batch_size: int):
        return (batch_size, old_shape[0] // batch_size) + tuple(old_shape[1:])

    def add_new_batch_dim(self, batch_size: int):
        if len(self) % batch_size != 0:
            raise ValueError("Current batch of transitions size is not a multiple of the new batch size. ")
        # the reference reshapes next_obs with obs's shape (types.py:72); same thing for valid batches
        shapes = {f: getattr(self, f).shape for f in _FIELDS}
        shapes["batch_next_obs"] = shapes["batch_obs"]
        return InteractionBatch(*(getattr(self, f).reshape(self._get_new_shape(shapes[f], batch_size)) for f in _FIELDS))


ModelInput = Union[torch.Tensor, InteractionBatch]
