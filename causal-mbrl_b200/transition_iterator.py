"""Host batching mirroring cmrl/util/transition_iterator.py: TransitionIterator (:31-95),
BootstrapIterator (:98-148, with-replacement member indices [E,N]) and _consolidate_batches (:14-28).

RNG consumption matches the reference (one `permutation(N)` per epoch when shuffling; one
`choice(N, (E,N))` or E `permutation(N)` at construction), so a seeded generator reproduces the
reference's batches.  `device_plan()` exposes the index tables for the on-device gather used by the
fused training path.
"""
from typing import Optional, Sequence

import numpy as np

from .types import InteractionBatch


def _consolidate_batches(batches: Sequence[InteractionBatch]) -> InteractionBatch:
    first = batches[0]
    n = len(batches)
    obs = np.empty((n,) + first.batch_obs.shape, dtype=first.batch_obs.dtype)
    act = np.empty((n,) + first.batch_action.shape, dtype=first.batch_action.dtype)
    next_obs = np.empty((n,) + first.batch_obs.shape, dtype=first.batch_obs.dtype)
    rewards = np.empty((n,) + first.batch_reward.shape, dtype=np.float32)
    dones = np.empty((n,) + first.batch_done.shape, dtype=bool)
    for i, b in enumerate(batches):
        obs[i], act[i], next_obs[i], rewards[i], dones[i] = b.as_tuple()
    return InteractionBatch(obs, act, next_obs, rewards, dones)


class TransitionIterator:
    """Iterates over batches of transitions, optionally reshuffled at the start of every pass."""

    def __init__(self, transitions: InteractionBatch, batch_size: int, shuffle_each_epoch: bool = False,
                 rng: Optional[np.random.Generator] = None):  # fmt: skip
        self.transitions = transitions
        self.num_stored = len(transitions)
        self._order: np.ndarray = np.arange(self.num_stored)
        self.batch_size = batch_size
        self._current_batch = 0
        self._shuffle_each_epoch = shuffle_each_epoch
        self._rng = rng if rng is not None else np.random.default_rng()

    def _get_indices_next_batch(self):
        start = self._current_batch * self.batch_size
        if start >= self.num_stored:
            raise StopIteration
        stop = min(start + self.batch_size, self.num_stored)
        self._current_batch += 1
        return self._order[start:stop]

    def __iter__(self):
        self._current_batch = 0
        if self._shuffle_each_epoch:
            self._order = self._rng.permutation(self.num_stored)
        return self

    def __next__(self):
        return self[self._get_indices_next_batch()]

    def ensemble_size(self):
        return 0

    def __len__(self):
        return (self.num_stored - 1) // self.batch_size + 1

    def __getitem__(self, item):
        return self.transitions[item]


class BootstrapIterator(TransitionIterator):
    """Each ensemble member sees its own bootstrap resample: batches are [E, B, ...]."""

    def __init__(self, transitions: InteractionBatch, batch_size: int, ensemble_size: int,
                 shuffle_each_epoch: bool = False, permute_indices: bool = True,
                 rng: Optional[np.random.Generator] = None):  # fmt: skip
        super().__init__(transitions, batch_size, shuffle_each_epoch=shuffle_each_epoch, rng=rng)
        self._ensemble_size = ensemble_size
        self._permute_indices = permute_indices
        self._bootstrap_iter = ensemble_size > 1
        self.member_indices = self._sample_member_indices()

    def _sample_member_indices(self) -> np.ndarray:
        if self._permute_indices:
            member_indices = np.empty((self.ensemble_size, self.num_stored), dtype=int)
            for i in range(self.ensemble_size):
                member_indices[i] = self._rng.permutation(self.num_stored)
            return member_indices
        return self._rng.choice(self.num_stored, size=(self.ensemble_size, self.num_stored), replace=True)

    def __iter__(self):
        super().__iter__()
        return self

    def __next__(self):
        if not self._bootstrap_iter:
            return super().__next__()
        indices = self._get_indices_next_batch()
        return _consolidate_batches([self[member_idx[indices]] for member_idx in self.member_indices])

    def toggle_bootstrap(self):
        """Toggles whether the iterator returns a batch per model or a single batch."""
        if self.ensemble_size > 1:
            self._bootstrap_iter = not self._bootstrap_iter

    @property
    def ensemble_size(self):
        return self._ensemble_size

    def device_plan(self):
        """(member_indices [E,N] int64, order [N] int64) of the *current* epoch: batch k of member e is
        transitions[member_indices[e, order[k*B:(k+1)*B]]] -- what the on-device gather replays."""
        return self.member_indices, self._order
