"""Transition-batch container and callback type aliases of the hot path.

API mirror of cmrl/types.py:20-77 (`InteractionBatch`: five parallel arrays, indexable, reshaped to `[n_batches, B, ...]`
by `add_new_batch_dim`); the methods are written once over the dataclass fields instead of per field.
"""
import dataclasses
from typing import Callable, Optional, Tuple, Union

import numpy as np
import torch

TensorType = Union[torch.Tensor, np.ndarray]
_Tensor3 = [torch.Tensor, torch.Tensor, torch.Tensor]  # (next_obs, pre_obs, action)
RewardFnType = Callable[_Tensor3, torch.Tensor]
TermFnType = Callable[_Tensor3, torch.Tensor]
InitObsFnType = Callable[[int], torch.Tensor]
ObsProcessFnType = Callable[[np.ndarray], np.ndarray]
TrajectoryEvalFnType = Callable[[TensorType, torch.Tensor], torch.Tensor]
InteractionData = Tuple[TensorType, ...]  # (obs, action, next_obs, reward, done)


class _ParallelArrays:
    """Behaviour shared by dataclasses whose fields are arrays with a common leading dimension."""

    @classmethod
    def _names(cls):
        return [f.name for f in dataclasses.fields(cls)]

    @property
    def attrs(self):
        return self._names()

    def _map(self, fn):
        return type(self)(*(fn(name, getattr(self, name)) for name in self._names()))

    def __len__(self):
        return getattr(self, self._names()[0]).shape[0]

    def as_tuple(self) -> InteractionData:
        return tuple(getattr(self, name) for name in self._names())

    def __getitem__(self, item):
        return self._map(lambda _, x: x[item])

    @staticmethod
    def _get_new_shape(old_shape: Tuple[int, ...], batch_size: int):
        return (batch_size, old_shape[0] // batch_size) + tuple(old_shape[1:])

    def add_new_batch_dim(self, batch_size: int):
        if len(self) % batch_size:
            raise ValueError("Current batch of transitions size is not a multiple of the new batch size. ")
        # quirk kept from the reference (types.py:72): next_obs is reshaped with the shape of obs
        shape_of = {name: getattr(self, name).shape for name in self._names()}
        if "batch_next_obs" in shape_of:
            shape_of["batch_next_obs"] = shape_of["batch_obs"]
        return self._map(lambda name, x: x.reshape(self._get_new_shape(shape_of[name], batch_size)))


@dataclasses.dataclass
class InteractionBatch(_ParallelArrays):
    """A batch of (obs, action, next_obs, reward, done) transitions."""

    batch_obs: Optional[TensorType]
    batch_action: Optional[TensorType]
    batch_next_obs: Optional[TensorType]
    batch_reward: Optional[TensorType]
    batch_done: Optional[TensorType]


ModelInput = Union[torch.Tensor, InteractionBatch]
