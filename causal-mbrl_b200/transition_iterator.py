"""Host-side batching with the interface of cmrl/util/transition_iterator.py: `TransitionIterator` (:31-95),
`BootstrapIterator` (:98-148: every member iterates its own resample of the data, batches are [E,B,...]) and
`_consolidate_batches` (:14-28).

RNG consumption matches the reference -- one `permutation(N)` per pass when shuffling; at construction E
`permutation(N)` calls, or one `choice(N, (E,N))` with replacement -- so a seeded generator reproduces the
reference's batches (tests/test_host_logic.py).  `device_plan()` exposes the index tables for the on-device gather of
the fused training path (`cmrl_bootstrap_gather`).
"""
from typing import Optional, Sequence

import numpy as np

from .types import InteractionBatch

# dtype the stacked field is stored in (None: that of the member batches; next_obs follows obs like the reference)
_STACK_DTYPE = {"batch_reward": np.float32, "batch_done": np.bool_}


def _consolidate_batches(batches: Sequence[InteractionBatch]) -> InteractionBatch:
    """[B,...] batches of the E members -> one batch of [E,B,...] arrays."""
    columns = []
    for name in batches[0].attrs:
        like = "batch_obs" if name == "batch_next_obs" else name
        dtype = _STACK_DTYPE.get(name, getattr(batches[0], like).dtype)
        columns.append(np.stack([getattr(b, name) for b in batches]).astype(dtype, copy=False))
    return InteractionBatch(*columns)


class TransitionIterator:
    """Passes over the transitions in batches of `batch_size` (the last one may be short); with
    `shuffle_each_epoch` every pass starts with a fresh permutation."""

    def __init__(self, transitions: InteractionBatch, batch_size: int, shuffle_each_epoch: bool = False,
                 rng: Optional[np.random.Generator] = None):  # fmt: skip
        self.transitions = transitions
        self.batch_size = batch_size
        self.num_stored = len(transitions)
        self._shuffle_each_epoch = shuffle_each_epoch
        self._owns_rng = rng is None  # a private generator: nobody else draws from it between two passes
        self._rng = np.random.default_rng() if rng is None else rng
        self._order: np.ndarray = np.arange(self.num_stored)  # row visited at each position of the pass
        self._next_order: Optional[np.ndarray] = None  # the next pass's permutation, drawn ahead of time
        self._current_batch = 0

    def __len__(self):
        return -(-self.num_stored // self.batch_size) if self.num_stored else 0

    def __getitem__(self, item):
        return self.transitions[item]

    def ensemble_size(self):
        return 0

    def _get_indices_next_batch(self):
        lo = self._current_batch * self.batch_size
        if lo >= self.num_stored:
            raise StopIteration
        self._current_batch += 1
        return self._order[lo:lo + self.batch_size]

    def __iter__(self):
        if self._shuffle_each_epoch:
            self._join_prefetch()
            if self._next_order is not None:
                self._order, self._next_order = self._next_order, None
            else:
                self._order = self._rng.permutation(self.num_stored)
        self._current_batch = 0
        return self

    def _join_prefetch(self):
        th = self.__dict__.pop("_prefetch_thread", None)
        if th is not None:
            th.join()

    def prefetch_order(self, background: bool = False, then=None) -> Optional[np.ndarray]:
        """Draw the NEXT pass's permutation ahead of time (the device-resident epoch does this while the GPU executes the
        current pass: at ~10 ns per row the shuffle of a large set otherwise sits between two epochs -- 80 ms for the
        8 M rows of config C4).  The generator is consumed in the same order as without the call -- the next `__iter__`
        uses exactly this draw -- but only a private generator is touched early: with a caller-supplied one another
        consumer could sit between the two draws.  background=True draws on a helper thread (numpy shuffles without the
        GIL) that the next `__iter__` / `prefetch_order()` joins; `then(order)` runs on that thread after the draw."""
        if not (self._shuffle_each_epoch and self._owns_rng):
            return None
        self._join_prefetch()
        if self._next_order is None:
            def draw():
                self._next_order = self._rng.permutation(self.num_stored)
                if then is not None:
                    then(self._next_order)

            if background:
                import threading

                th = self.__dict__["_prefetch_thread"] = threading.Thread(target=draw, daemon=True)
                th.start()
                return None
            draw()
        return self._next_order

    def __next__(self):
        return self[self._get_indices_next_batch()]


class BootstrapIterator(TransitionIterator):
    """Each ensemble member sees its own resample of the transitions (`member_indices` [E,N]: a permutation per
    member, or N draws with replacement when `permute_indices` is off); a batch is the same positions of every
    member's resample, stacked to [E,B,...].  `toggle_bootstrap()` switches to plain batches and back."""

    def __init__(self, transitions: InteractionBatch, batch_size: int, ensemble_size: int,
                 shuffle_each_epoch: bool = False, permute_indices: bool = True,
                 rng: Optional[np.random.Generator] = None):  # fmt: skip
        super().__init__(transitions, batch_size, shuffle_each_epoch=shuffle_each_epoch, rng=rng)
        self._ensemble_size = ensemble_size
        self._permute_indices = permute_indices
        self._bootstrap_iter = ensemble_size > 1
        self.member_indices = self._sample_member_indices()

    @property
    def ensemble_size(self):
        return self._ensemble_size

    def _sample_member_indices(self) -> np.ndarray:
        E, N = self._ensemble_size, self.num_stored
        if not self._permute_indices:
            return self._rng.choice(N, size=(E, N), replace=True)
        return np.stack([self._rng.permutation(N) for _ in range(E)]).astype(int, copy=False).reshape(E, N)

    def toggle_bootstrap(self):
        if self._ensemble_size > 1:
            self._bootstrap_iter = not self._bootstrap_iter

    def __next__(self):
        if not self._bootstrap_iter:
            return super().__next__()
        positions = self._get_indices_next_batch()
        return _consolidate_batches([self[rows[positions]] for rows in self.member_indices])

    def device_plan(self):
        """(member_indices [E,N] int64, order [N] int64) of the *current* epoch: batch k of member e is
        transitions[member_indices[e, order[k*B:(k+1)*B]]] -- what the on-device gather replays."""
        return self.member_indices, self._order
