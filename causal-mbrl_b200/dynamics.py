"""BaseDynamics / PlainEnsembleDynamics / ConstraintBasedDynamics with the reference's public interface
(cmrl/models/dynamics/base_dynamics.py:22-221, plain_dynamics.py:18-142, constraint_based_dynamics.py:22-183);
the arithmetic runs on libcmrl_b200 kernels:

* `train`    -- per batch: K1 forward (saving pre-activations), K3 NLL forward + backward, K2 backward, K1
                backward writing straight into the flat gradient arena, one fused L2-Adam launch
                (no autograd graph, no per-parameter optimiser loop).
* `evaluate` -- K1/K2 forward on un-replicated rows + K3 MSE; `evaluate_member_mse` reduces on device to
                the [E] vector elite selection needs.
* `query`    -- all-member forward on un-replicated rows (the reference repeats the input E times).

Deliberate differences from the reference: a learned termination mech trains on `batch_done`
(the reference raises KeyError 'batch_terminal').
"""
import os
import abc
import collections
import copy
import itertools
import pathlib
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from . import _lib, ops
from .mlp import EnsembleMLP
from .optim import FusedAdamL2
from .reward_mech import BaseRewardMech
from .termination_mech import BaseTerminationMech
from .transition import BaseTransition
from .transition_iterator import BootstrapIterator, TransitionIterator
from .types import InteractionBatch
from .util import to_tensor


def dp_row_slice(num_rows: int, world_size: int, rank: int) -> slice:
    """Rows of an ensemble batch owned by `rank` in data-parallel training: contiguous, remainder to the low
    ranks, so ragged last batches are covered exactly once."""
    base, rem = divmod(num_rows, world_size)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def net_shard_range(num_subnets: int, world_size: int, rank: int):
    """Output dimensions (sub-networks) of rank `rank` when `num_subnets` are dealt to `world_size` ranks in contiguous,
    near-equal shares (the first `num_subnets % world_size` ranks take one more)."""
    sl = dp_row_slice(num_subnets, world_size, rank)
    return sl.start, sl.stop


def dp_allreduce_member_sums(sums: torch.Tensor, group) -> torch.Tensor:
    """[E] per-member sums over one rank's row shard -> the global sums, identical on every rank (sharded
    evaluation; every rank then derives the same validation scores and the same elite set)."""
    import torch.distributed as dist

    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


class BaseDynamics:
    _MECH_TO_VARIABLE = {
        "transition": "batch_next_obs",
        "reward_mech": "batch_reward",
        "termination_mech": "batch_terminal",
    }
    _VARIABLE_TO_MECH = dict([(value, key) for key, value in _MECH_TO_VARIABLE.items()])

    def __init__(
        self,
        transition: BaseTransition,
        learned_reward: bool = True,
        reward_mech: Optional[BaseRewardMech] = None,
        learned_termination: bool = False,
        termination_mech: Optional[BaseTerminationMech] = None,
        optim_lr: float = 1e-4,
        weight_decay: float = 1e-5,
        optim_eps: float = 1e-8,
        logger=None,
    ):
        super(BaseDynamics, self).__init__()
        self.transition = transition
        self.learned_reward = learned_reward
        self.reward_mech = reward_mech
        self.learned_termination = learned_termination
        self.termination_mech = termination_mech

        self.optim_lr = optim_lr
        self.weight_decay = weight_decay
        self.optim_eps = optim_eps
        self.logger = logger

        self.device = self.transition.device
        self.ensemble_num = self.transition.ensemble_num

        self.learn_mech = ["transition"]
        self.transition_optimizer = self._make_optimizer(self.transition)
        if self.learned_reward:
            self.reward_mech_optimizer = self._make_optimizer(self.reward_mech)
            self.learn_mech.append("reward_mech")
        if self.learned_termination:
            self.termination_mech_optimizer = self._make_optimizer(self.termination_mech)
            self.learn_mech.append("termination_mech")

        self.total_epoch = {}
        for mech in self.learn_mech:
            self.total_epoch[mech] = 0

        # data-parallel training (optional): torch.distributed process group + this rank's row slice
        self.dp_group = None
        self.overlap_dw = True   # train_step: weight-gradient GEMMs on a side stream next to the dz chain
        self._side_stream = None

    def _make_optimizer(self, model):
        return FusedAdamL2(model, lr=self.optim_lr, weight_decay=self.weight_decay, eps=self.optim_eps)

    @abc.abstractmethod
    def learn(self, replay_buffer, **kwargs):
        pass

    # ------------------------------------------------------------------ "single batch data" helpers
    def get_3d_tensor(self, data: Union[np.ndarray, torch.Tensor], is_ensemble: bool):
        """[E,B,D] device tensor.  Non-ensemble data is broadcast over members as a stride-0 view instead of
        the reference's physical `repeat([E,1,1])` (base_dynamics.py:90-100)."""
        if isinstance(data, np.ndarray):
            data = torch.from_numpy(data)
        if data.dtype != torch.float32:
            data = data.to(torch.float32)
        if is_ensemble:
            if data.ndim == 2:  # reward or terminal
                data = data.unsqueeze(data.ndim)
            return data.to(self.device)
        if data.ndim == 1:  # reward or terminal
            data = data.unsqueeze(data.ndim)
        return data.to(self.device).unsqueeze(0).expand(self.ensemble_num, -1, -1)

    def _batch_variable(self, batch: InteractionBatch, variable: str):
        if variable == "batch_terminal":  # reference quirk 1: InteractionBatch calls it batch_done
            return batch.batch_done
        return getattr(batch, variable)

    def _mech_target(self, batch: InteractionBatch, mech: str):
        """The variable `mech` is trained on (base_dynamics.py:27-28); a net-sharded transition takes its own columns."""
        return getattr(self, mech).target_columns(self._batch_variable(batch, self._MECH_TO_VARIABLE[mech]))

    def get_mech_loss(self, batch: InteractionBatch, mech: str = "transition", loss_type: str = "default",
                      is_ensemble: bool = False):  # fmt: skip
        model = getattr(self, mech)
        model_in = {
            "batch_obs": self.get_3d_tensor(batch.batch_obs, is_ensemble=is_ensemble),
            "batch_action": self.get_3d_tensor(batch.batch_action, is_ensemble=is_ensemble),
        }
        if loss_type == "default":
            loss_type = "mse" if model.deterministic else "nll"
        target = self.get_3d_tensor(self._mech_target(batch, mech), is_ensemble=is_ensemble)
        get_loss = getattr(model, "get_{}_loss".format(loss_type))
        return get_loss(model_in, target)

    # ------------------------------------------------------------------ "replay buffer" helper
    def dataset_split(self, replay_buffer, validation_ratio: float = 0.2, batch_size: int = 256,
                      shuffle_each_epoch: bool = True, bootstrap_permutes: bool = False,
                      ) -> Tuple[TransitionIterator, Optional[TransitionIterator]]:  # fmt: skip
        size = replay_buffer.buffer_size if replay_buffer.full else replay_buffer.pos
        data = InteractionBatch(
            replay_buffer.observations[:size, 0].astype(np.float32),
            replay_buffer.actions[:size, 0],
            replay_buffer.next_observations[:size, 0].astype(np.float32),
            replay_buffer.rewards[:size, 0],
            replay_buffer.dones[:size, 0],
        )
        val_size = int(len(data) * validation_ratio)
        train_size = len(data) - val_size
        train_iter = BootstrapIterator(data[:train_size], batch_size, self.ensemble_num,
                                       shuffle_each_epoch=shuffle_each_epoch, permute_indices=bootstrap_permutes)  # fmt: skip
        val_iter = None
        if val_size > 0:
            val_iter = TransitionIterator(data[train_size:], batch_size, shuffle_each_epoch=False)
        return train_iter, val_iter

    # ------------------------------------------------------------------ "dataset" helpers
    def evaluate(self, dataset: TransitionIterator, mech: str = "transition"):
        """Hold-out MSE, un-reduced: [E, N_val, D] on the CPU like base_dynamics.py:159-171."""
        assert not isinstance(dataset, BootstrapIterator)
        batch_loss_list = []
        with torch.no_grad():
            for batch in dataset:
                batch_loss_list.append(self.get_mech_loss(batch, mech=mech, loss_type="mse", is_ensemble=False))
        return torch.cat(batch_loss_list, dim=batch_loss_list[0].ndim - 2).cpu()

    def evaluate_member_mse(self, dataset: TransitionIterator, mech: str = "transition", chunk: int = 65536):
        """`evaluate(...).mean(dim=(1, 2))` -> [E] without materialising the [E,N,D] loss: the whole hold-out
        set goes through the forward kernels in large chunks and K3 reduces per member on the device (fp64)."""
        assert not isinstance(dataset, BootstrapIterator)
        model = getattr(self, mech)
        data = dataset.transitions
        target_np = self._mech_target(data, mech)
        sums = torch.zeros(self.ensemble_num, device=self.device, dtype=torch.float64)
        n, d = len(data), model._head_dim
        # data parallel: every rank evaluates its contiguous share of the hold-out rows, one [E] all-reduce at the end
        rows = dp_row_slice(n, self.dp_world, self.dp_rank) if self.dp_group is not None else slice(0, n)
        # every member's mean of every hold-out row: one pass of the fused tcgen05 kernel (3xTF32, fp32-grade) over all-member
        # buckets where it takes the model, else the all-member forward of the per-layer kernels
        pack = model.packed_tf32() if (hasattr(model, "hidden_layers") and not model.deterministic
                                       and torch.device(self.device).type == "cuda") else None  # (wrappers such as ForwardEuler: forward)
        buckets = self.__dict__.setdefault("_eval_buckets", {})
        with torch.no_grad():
            for s in range(rows.start, rows.stop, chunk):
                sl = slice(s, min(rows.stop, s + chunk))
                obs = self.get_3d_tensor(data.batch_obs[sl], is_ensemble=False)
                act = self.get_3d_tensor(data.batch_action[sl], is_ensemble=False)
                target = self.get_3d_tensor(target_np[sl], is_ensemble=False)
                if pack is not None and pack.ok:
                    E, nr = self.ensemble_num, obs.shape[1]
                    if nr not in buckets:
                        if len(buckets) >= 4:
                            buckets.clear()
                        buckets[nr] = (torch.arange(nr, device=obs.device, dtype=torch.int32).repeat(E),
                                       (torch.arange(E + 1, device=obs.device, dtype=torch.int64) * nr).to(torch.int32))
                    perm, offsets = buckets[nr]
                    mn, mx = model._bounds()
                    mean = ops.mlp_rollout_tf32(pack, obs[0].contiguous(), act[0].contiguous(), perm, offsets, model._act_kind,
                                                model._residual(), mn, mx, None, all_members=True)  # fmt: skip
                else:
                    mean, _ = model.forward(obs, act)
                d = mean.shape[-1]
                ops.mse_member_sums(mean, target, sums)
        if self.dp_group is not None:
            dp_allreduce_member_sums(sums, self.dp_group)
        ns_group = getattr(self, "ns_group", None)
        if ns_group is not None and getattr(model, "is_net_shard", False):
            # net sharding: every rank scored its own output dimensions; the [E] sums of all shards make the mechanism's score
            dp_allreduce_member_sums(sums, ns_group)
            d = model._grad_dims()
        return (sums / float(n * d)).to(torch.float32).cpu()

    def train_step(self, mech: str, obs: torch.Tensor, act: torch.Tensor, target: torch.Tensor,
                   want_loss: bool = True, grad_rows: Optional[int] = None, graphable: bool = False,
                   data_ready=None, loss_ready=None):  # fmt: skip
        """One optimiser step on an ensemble batch (obs/act/target: [E,B,*] device tensors), equivalent to
        `loss = get_nll_loss(...); zero_grad(); loss.mean().backward(); step()` (base_dynamics.py:181-186).
        Returns the un-reduced loss [E,B,D] (None if not wanted).  grad_rows: global rows per member when the
        batch is one shard of a data-parallel step.  data_ready: called (once) right before the first kernel that reads
        the batch -- the device-resident epoch assembles the batch on a side stream next to the weight packing and joins
        the streams there.  loss_ready(loss): called by the fused step as soon as the loss is enqueued (the epoch's loss
        bookkeeping then runs on a side stream under the weight-gradient kernel); returns True there, else it is not called."""
        model = getattr(self, mech)
        optim = getattr(self, "{}_optimizer".format(mech))
        if data_ready is not None and ((model.deterministic or any(p.requires_grad for p in model._bounds() if p is not None)
                                        or not hasattr(model, "hidden_layers")) or obs.shape[1] == 0 or obs.dim() != 3
                                       or ops.fused_train_plan(model, obs.shape[0], obs.shape[1]) is None):
            data_ready()  # only the fused plan has work (the packing) that does not depend on the batch
            data_ready = None
        if obs.shape[1] == 0:
            # data parallel, ragged last batch with fewer rows than ranks: this rank has no rows, but it still has to
            # contribute (zero) gradients to the all-reduce and apply the same Adam update as every other rank
            _, flat_grad = model.flat_parameters()
            flat_grad.zero_()
            self._sync_and_step(model, optim, graphable)
            return torch.empty((obs.shape[0], 0, target.shape[-1]), device=obs.device, dtype=torch.float32) if want_loss else None
        if (model.deterministic or any(p.requires_grad for p in model._bounds() if p is not None)
                or not hasattr(model, "hidden_layers")):
            # uncommon configurations (deterministic heads, learned logvar bounds, wrappers without a layer stack of
            # their own such as ForwardEulerTransition) go through the autograd nodes (same kernels)
            loss = (model.get_mse_loss if model.deterministic else model.get_nll_loss)(
                {"batch_obs": obs, "batch_action": act}, target)  # fmt: skip
            optim.zero_grad()
            (loss.mean() * (model._head_dim / float(model._grad_dims()))).backward()
            self._sync_grads(model)
            optim.step()
            return loss.detach()
        E, B = obs.shape[0], obs.shape[1]
        flat, flat_grad = model.flat_parameters()
        plan = ops.fused_train_plan(model, E, B) if obs.dim() == 3 else None
        if plan is not None:
            # forward, Gaussian head, NLL, dz chain and every weight gradient on tcgen05 (3xTF32): three launches
            mn, mx = model._bounds()
            rows = B if grad_rows is None else grad_rows
            loss = plan.step(obs, act, target, mn, mx, model._nll_reg(), 1.0 / float(E * rows * model._head_dim), want_loss,
                             after_pack=data_ready, after_fused=loss_ready)
            self._sync_and_step(model, optim, graphable)
            return loss
        params = model._stack_params()
        L = model.num_layers
        kind = model._act_kind
        mask = model._kernel_mask()
        mask_str = ops.mask_strides(mask, model._parallel_num(), E, B) if mask is not None else (0, 0, 0)
        obs_c, act_c = obs.contiguous(), act.contiguous()
        x = ops.gather_inputs(obs_c.view(E * B, -1), act_c.view(E * B, -1), None).view(E, B, -1)
        hs, zs, h = [x], [], x
        for i in range(L):
            h, z = ops.ens_linear_fwd(h, params[2 * i], params[2 * i + 1], kind, mask if i == 0 else None, mask_str,
                                      want_z=True)  # fmt: skip
            hs.append(h)
            zs.append(z)
        raw = ops.ens_linear_fwd(h, params[2 * L], params[2 * L + 1], ops.ACT_IDENTITY)
        mn, mx = model._bounds()
        D = model._head_dim
        mean, logvar = model._head_fwd(raw, obs_c if model._residual() else None, E, B)
        loss = ops.nll_loss_fwd(mean, logvar, target, model._nll_reg()) if want_loss else None
        rows = B if grad_rows is None else grad_rows
        g_mean, g_logvar = ops.nll_loss_bwd(mean, logvar, target, g_scalar=1.0 / float(E * rows * model._grad_dims()))
        dz = ops.gaussian_head_bwd(raw, model._head_layout, E, B, D, mn, mx, g_mean, g_logvar)
        # Backward: dW_i and dx_i both need only dz_i, so the weight-gradient GEMMs run on a side stream next to the
        # dz chain (at the reference batch size every GEMM is a latency-bound ~17 us launch on < 148 CTAs: the critical
        # path shrinks from 2L+1 to L+1 GEMMs).  Works under CUDA-graph capture (fork / join through events).
        main = torch.cuda.current_stream()
        side = self._side_stream if self.overlap_dw else None
        if side is None and self.overlap_dw:
            side = self._side_stream = torch.cuda.Stream(device=obs.device)
        keep = []  # dz tensors stay referenced until the join: the side stream reads them
        for i in range(L, -1, -1):
            w, b = params[2 * i], params[2 * i + 1]
            h_in = hs[i]
            if i == 0 and h_in.dim() < w.dim():
                h_in = h_in.expand(tuple(w.shape[:-2]) + tuple(h_in.shape[-2:]))
            if side is not None:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    ops.ens_linear_bwd_dw(h_in, dz, w.shape, mask if i == 0 else None, mask_str, dw=w.grad, db=b.grad)
                keep.append(dz)
            else:
                ops.ens_linear_bwd_dw(h_in, dz, w.shape, mask if i == 0 else None, mask_str, dw=w.grad, db=b.grad)
            if i > 0:
                dz = ops.ens_linear_bwd_dx(dz, w, zs[i - 1], kind)
        if side is not None:
            main.wait_stream(side)
        del keep
        self._sync_and_step(model, optim, graphable)
        return loss

    def enable_data_parallel(self, group=None, sync: str = "auto"):
        """Shard every ensemble batch's rows over the ranks of `group` (default: the world group) in `train`;
        gradients are summed over the flat arena and the loss scale uses the global row count, so every rank applies
        the identical Adam update.  sync="peer": ONE kernel per step exchanges the gradients over NVLink
        peer memory and applies Adam (`cmrl_allreduce_grads_adam`, csrc/dp_sync.cu); sync="nccl": one
        `torch.distributed.all_reduce` followed by the Adam kernel; "auto" = peer for CUDA models on an NCCL group of
        more than one rank (every rank must make this call: the peer set-up is collective)."""
        import torch.distributed as dist

        assert sync in ("auto", "peer", "nccl")
        self.dp_group = group if group is not None else dist.group.WORLD
        self.dp_world = dist.get_world_size(self.dp_group)
        self.dp_rank = dist.get_rank(self.dp_group)
        self._peer_sync = None
        on_cuda = torch.device(self.device).type == "cuda"
        if sync == "auto" and os.environ.get("CMRL_DP_SYNC") in ("peer", "nccl"):  # operator override (A/B timing)
            sync = os.environ["CMRL_DP_SYNC"]
        auto = sync == "auto"
        if auto:
            sync = "peer" if (on_cuda and self.dp_world > 1 and dist.get_backend(self.dp_group) == "nccl") else "nccl"
        if sync == "peer":
            from .peer_sync import PeerGradSync

            assert on_cuda, "the peer-memory gradient exchange needs the models on a CUDA device"
            largest = 1
            for mech in self.learn_mech:
                model = getattr(self, mech)
                model.flat_parameters()
                largest = max(largest, int(model._n_trainable))
            try:
                self._peer_sync = PeerGradSync(self.dp_group, largest, self.device)
            except ops._lib.CmrlB200Error:
                if not auto:
                    raise
                self._peer_sync = None  # consistent on all ranks (the set-up ends with an all-reduce of its outcome): NCCL path

    def enable_net_sharding(self, group=None):
        """The transition of this dynamics is one rank's NET SHARD of an ExternalMaskTransition (`subnet_range=`, see
        `net_shard_range`): training needs no exchange at all -- the sub-networks are independent -- and the hold-out score
        that drives early stopping and elite selection is the sum of every shard's [E] sums (one all-reduce per evaluation),
        so all ranks stop at the same epoch and pick the same elites.  `gather_net_shards()` then assembles the full model."""
        import torch.distributed as dist

        assert getattr(self.transition, "is_net_shard", False), "build the transition with subnet_range=net_shard_range(...)"
        self.ns_group = group if group is not None else dist.group.WORLD

    def gather_net_shards(self):
        """Replace the shard by the full transition (collective), e.g. before rolling it out in a VecFakeEnv."""
        full = self.transition.gather_from_shards(self.ns_group)
        self.transition = full
        self.ns_group = None
        self.transition_optimizer = self._make_optimizer(full)
        return full

    def broadcast_parameters(self, src: int = 0, group=None):
        """Rollout replicas: envs are sharded over ranks with a full weight replica each and no per-step
        collective, so after `learn` ran on one rank (or a checkpoint was loaded there) every mechanism's parameters
        (one broadcast of the flat arena) and elite set are copied from global rank `src`.  Not needed after
        data-parallel training, which leaves identical replicas."""
        import torch.distributed as dist

        if group is None:
            group = self.dp_group if self.dp_group is not None else dist.group.WORLD
        for mech in self.learn_mech:
            model = getattr(self, mech)
            try:
                tensors = [model.flat_parameters()[0]]
            except ops._lib.CmrlB200Error:  # parameters not on a CUDA device (host-logic tests): one broadcast each
                tensors = [p.data for p in model.parameters()]
            for x in tensors:
                dist.broadcast(x, src=src, group=group)
            elite = torch.tensor([int(e) for e in model.elite_members], dtype=torch.int64, device=tensors[0].device)
            dist.broadcast(elite, src=src, group=group)
            model.set_elite_members([int(e) for e in elite.tolist()])

    def _sync_grads(self, model):
        """Data-parallel gradient exchange alone (autograd paths): the peer-memory kernel or one NCCL all-reduce (sum)
        over the flat gradient arena."""
        if self.dp_group is not None:
            _, flat_grad = model.flat_parameters()
            peer = getattr(self, "_peer_sync", None)
            if peer is not None:
                peer.allreduce(flat_grad[: model._n_trainable])
                return
            import torch.distributed as dist

            dist.all_reduce(flat_grad[: model._n_trainable], op=dist.ReduceOp.SUM, group=self.dp_group)

    def _sync_and_step(self, model, optim, graphable):
        """Gradient exchange + optimiser step: fused into one launch on the peer-memory path."""
        peer = getattr(self, "_peer_sync", None) if self.dp_group is not None else None
        if peer is not None:
            optim.step_synced(peer)
            return
        self._sync_grads(model)
        if graphable:
            optim.step_graphable()
        else:
            optim.step()

    def train_epoch_device(self, dataset: BootstrapIterator, mech: str = "transition", use_graph: bool = True):
        """One pass over the bootstrap iterator with the training set, the bootstrap index table and the epoch order
        resident on the device: batches are assembled by `cmrl_bootstrap_gather` and the whole optimiser step
        (forward, NLL, backward, Adam; ~20 launches) is captured once in a CUDA graph and replayed per batch.
        Consumes the iterator's RNG exactly like `train` (one permutation per epoch) and performs the same updates;
        returns the mean train loss (what `learn` logs) instead of the [E,N,D] tensor."""
        assert isinstance(dataset, BootstrapIterator)
        model = getattr(self, mech)
        optim = getattr(self, "{}_optimizer".format(mech))
        fast = (not model.deterministic and dataset._bootstrap_iter and hasattr(model, "hidden_layers")
                and not any(q.requires_grad for q in model._bounds()))  # fmt: skip
        if not fast:
            return float(self.train(dataset, mech=mech).mean().item())
        runner = self._epoch_runners.get(mech) if hasattr(self, "_epoch_runners") else None
        if not hasattr(self, "_epoch_runners"):
            self._epoch_runners = {}
        if runner is None or runner.dataset is not dataset:
            runner = _DeviceEpochRunner(self, mech, dataset, use_graph)
            self._epoch_runners[mech] = runner
        return runner.run_epoch()

    def train(self, dataset: TransitionIterator, mech: str = "transition"):
        """One pass over the bootstrap iterator; returns the un-reduced train loss [E, N_train, D] on the CPU
        like base_dynamics.py:173-188."""
        assert isinstance(dataset, BootstrapIterator)
        variable = self._MECH_TO_VARIABLE[mech]
        batch_loss_list = []
        for batch in dataset:
            obs = self.get_3d_tensor(batch.batch_obs, is_ensemble=True)
            act = self.get_3d_tensor(batch.batch_action, is_ensemble=True)
            target = self.get_3d_tensor(self._batch_variable(batch, variable), is_ensemble=True)
            if self.dp_group is not None:
                rows = obs.shape[1]
                sl = dp_row_slice(rows, self.dp_world, self.dp_rank)
                batch_loss_list.append(self.train_step(mech, obs[:, sl], act[:, sl], target[:, sl], grad_rows=rows))
            else:
                batch_loss_list.append(self.train_step(mech, obs, act, target))
        return torch.cat(batch_loss_list, dim=batch_loss_list[0].ndim - 2).detach().cpu()

    def query(self, obs, action, return_as_np=True):
        """All-member prediction: variable -> {"mean", "logvar"} as [E,B,D] (base_dynamics.py:190-204)."""
        result = collections.defaultdict(dict)
        obs = self.get_3d_tensor(obs, is_ensemble=False)
        action = self.get_3d_tensor(action, is_ensemble=False)
        for mech in self.learn_mech:
            with torch.no_grad():
                mean, logvar = getattr(self, "{}".format(mech)).forward(obs, action)
            variable = self.get_variable_by_mech(mech)
            if return_as_np:
                result[variable]["mean"] = mean.cpu().numpy()
                result[variable]["logvar"] = logvar.cpu().numpy()
            else:
                result[variable]["mean"] = mean.cpu()
                result[variable]["logvar"] = logvar.cpu()
        return result

    # ------------------------------------------------------------------ other helpers
    def save(self, save_dir: Union[str, pathlib.Path]):
        for mech in self.learn_mech:
            getattr(self, mech).save(save_dir=save_dir)

    def load(self, load_dir: Union[str, pathlib.Path], load_device: Optional[str] = None):
        for mech in self.learn_mech:
            getattr(self, mech).load(load_dir=load_dir, load_device=load_device)

    def get_variable_by_mech(self, mech: str) -> str:
        assert mech in self._MECH_TO_VARIABLE
        return self._MECH_TO_VARIABLE[mech]

    def get_mach_by_variable(self, variable: str) -> str:
        assert variable in self._VARIABLE_TO_MECH
        return self._VARIABLE_TO_MECH[variable]


class _DeviceEpochRunner:
    """Device-resident training epoch: dataset + bootstrap tables in HBM, on-device batch gather,
    CUDA-graph replay of the optimiser step."""

    def __init__(self, dyn, mech, dataset, use_graph=True):
        self.dyn, self.mech, self.dataset, self.use_graph = dyn, mech, dataset, use_graph
        dev = dyn.device
        data = dataset.transitions
        target = dyn._mech_target(data, mech)

        def dev_f32(a):
            a = np.ascontiguousarray(a, dtype=np.float32)
            return torch.from_numpy(a.reshape(a.shape[0], -1)).to(dev)

        self.obs, self.act, self.target = dev_f32(data.batch_obs), dev_f32(data.batch_action), dev_f32(target)
        self.N = self.obs.shape[0]
        self.member_idx = torch.from_numpy(np.ascontiguousarray(dataset.member_indices, dtype=np.int32)).to(dev)
        self.order = torch.zeros(self.N, dtype=torch.int32, device=dev)
        self.start = torch.zeros(2, dtype=torch.int32, device=dev)  # [first row of the next batch, the gather kernel's arrival ticket]
        E, B = dataset.ensemble_size, dataset.batch_size
        self.E, self.B = E, B
        # data parallel (every rank holds the same set and epoch order): this rank gathers and trains only its row slice
        # of each batch; train_step all-reduces the flat gradient arena and scales the loss by the global row count
        self.row_slice = dp_row_slice(B, dyn.dp_world, dyn.dp_rank) if dyn.dp_group is not None else slice(0, B)
        Bl = self.Bl = self.row_slice.stop - self.row_slice.start
        self.b_obs = torch.empty((E, Bl, self.obs.shape[1]), device=dev)
        self.b_act = torch.empty((E, Bl, self.act.shape[1]), device=dev)
        self.b_tgt = torch.empty((E, Bl, self.target.shape[1]), device=dev)
        self.loss_sum = torch.zeros((), device=dev, dtype=torch.float64)
        self.sum_ws = torch.zeros(66, device=dev, dtype=torch.float64)
        self.side = torch.cuda.Stream(device=dev)
        self.graph = None
        self._graph_key = None
        self._order_pinned = [torch.empty(self.N, dtype=torch.int32).pin_memory() for _ in range(2)]
        self._order_flip = 0
        self._staged_order = None  # (the int64 permutation object, its int32 copy in pinned memory)

    def _stage_next(self, order):
        self._staged_order = (order, self._stage_order(order))

    def _stage_order(self, order):
        """int32 copy of an epoch order in one of two pinned buffers (the other may still be in flight to the device)."""
        buf = self._order_pinned[self._order_flip]
        self._order_flip ^= 1
        buf.numpy()[:] = order
        return buf

    def _step(self):
        # `start` points at this rank's first row of the batch and advances by the global batch size
        # the batch is assembled on a side stream while the main stream packs the weights (fork / join: also under capture)
        main = torch.cuda.current_stream()
        join = None
        if self.Bl > 0:
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                ops.bootstrap_gather(self.obs, self.act, self.target, self.member_idx, self.order, self.start, self.B, self.Bl,
                                     self.b_obs, self.b_act, self.b_tgt)  # fmt: skip
            join = lambda: main.wait_stream(self.side)  # noqa: E731
        summed = []

        def sum_on_side(loss):  # the running epoch loss, under the weight-gradient kernel
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                ops.sum_acc_f64(loss, self.loss_sum, self.sum_ws)
            summed.append(True)

        loss = self.dyn.train_step(self.mech, self.b_obs, self.b_act, self.b_tgt, want_loss=True, graphable=True,
                                   grad_rows=self.B if self.dyn.dp_group is not None else None, data_ready=join,
                                   loss_ready=sum_on_side)  # fmt: skip
        if summed:
            main.wait_stream(self.side)
        elif loss.numel() > 0:
            ops.sum_acc_f64(loss, self.loss_sum, self.sum_ws)

    def run_epoch(self):
        ds = self.dataset
        iter(ds)  # reshuffles exactly like the reference iterator (one rng.permutation per epoch); joins a background draw
        staged = self._staged_order
        if staged is None or staged[0] is not ds._order:
            staged = (ds._order, self._stage_order(ds._order))
        self._staged_order = None
        self.order.copy_(staged[1], non_blocking=True)
        # the NEXT epoch's permutation is drawn and staged on a helper thread while this epoch is launched and executed
        ds.prefetch_order(background=True, then=self._stage_next)
        self.start[:1].fill_(self.row_slice.start)
        self.loss_sum.zero_()
        model = getattr(self.dyn, self.mech)
        model._nll_reg()  # host-side constant, cached before any capture
        flat, flat_grad = model.flat_parameters()
        plan = None
        if self.Bl > 0:
            plan = ops.fused_train_plan(model, self.E, self.Bl)  # buffers of the fused tcgen05 step: allocated before the capture
            if plan is not None:
                plan.pinned = True  # the captured graph bakes their addresses
        optim = getattr(self.dyn, "{}_optimizer".format(self.mech))
        optim.prepare_graphable()
        n_full = self.N // self.B
        ws = None
        if self.Bl >= 2048:  # split-K dw workspace (ops.ens_linear_bwd_dw) must exist before the step is captured
            worst = max((p.numel() // p.shape[-2]) * (p.shape[-2] + 1) for p in model.parameters() if p.dim() >= 3 and p.shape[-2] > 1)
            ws = ops.splitk_workspace(self.obs.device, min(32 * worst, (256 << 20) // 4))
        # a captured step bakes the addresses of the arena, the Adam moments, the plan's buffers, the mask and the workspace:
        # when one of them was replaced since the capture (model.to(), set_input_mask, an optimiser state loaded from a
        # checkpoint, a larger workspace) the step is captured again
        mask = model._kernel_mask()
        graph_key = (flat.data_ptr(), flat_grad.data_ptr(), None if plan is None else (plan.packed.data_ptr(), plan.scratch.data_ptr()),
                     None if mask is None else (mask.data_ptr(), mask._version),
                     None if ws is None else ws.data_ptr(), optim._exp_avg.data_ptr(), optim._exp_avg_sq.data_ptr())  # fmt: skip
        if self.graph is not None and graph_key != self._graph_key:
            self.graph = None
        self._graph_key = graph_key
        if self.use_graph and self.graph is None and n_full > 0:
            if self.dyn.dp_group is not None and getattr(self.dyn, "_peer_sync", None) is None:
                self._step()  # one eager step first: NCCL sets its communicator up outside the capture
                n_full -= 1
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            l0 = _lib.launch_count()
            with ops.graph_capture(g):
                self._step()
            self.graph_launches = _lib.launch_count() - l0
            self.graph = g
        for _ in range(n_full):
            if self.graph is not None:
                self.graph.replay()
                _lib.note_replayed(self.graph_launches)
                # the captured Adam step ran again (no Python inside a replay)
                getattr(self.dyn, "{}_optimizer".format(self.mech))._mark_weights_changed()
            else:
                self._step()
        total = (self.N // self.B) * self.B
        rem = self.N - total
        if rem > 0:  # ragged last batch: same rows the reference iterator would yield
            dp = self.dyn.dp_group is not None
            sl = dp_row_slice(rem, self.dyn.dp_world, self.dyn.dp_rank) if dp else slice(0, rem)
            idx = torch.from_numpy(np.ascontiguousarray(ds._order[total:][sl], dtype=np.int64)).to(self.obs.device)
            rows = self.member_idx.long()[:, idx]  # [E, local rem]
            loss = self.dyn.train_step(self.mech, self.obs[rows], self.act[rows], self.target[rows], want_loss=True,
                                       graphable=True, grad_rows=rem if dp else None)  # fmt: skip
            if loss.numel() > 0:
                ops.sum_acc_f64(loss.contiguous(), self.loss_sum, self.sum_ws)
        ds._current_batch = len(ds)  # the iterator is exhausted, like after a `for batch in dataset` loop
        if self.dyn.dp_group is not None:
            import torch.distributed as dist

            dist.all_reduce(self.loss_sum, op=dist.ReduceOp.SUM, group=self.dyn.dp_group)
            if getattr(self.dyn, "_peer_sync", None) is not None:
                self.dyn._peer_sync.check()
        denom = float(self.E * self.N * self.b_tgt.shape[-1])
        return float(self.loss_sum.item()) / denom


class PlainEnsembleDynamics(BaseDynamics):
    def __init__(
        self,
        transition: BaseTransition,
        learned_reward: bool = True,
        reward_mech: Optional[BaseRewardMech] = None,
        learned_termination: bool = False,
        termination_mech: Optional[BaseTerminationMech] = None,
        # trainer
        optim_lr: float = 1e-4,
        weight_decay: float = 1e-5,
        optim_eps: float = 1e-8,
        logger=None,
    ):
        super(PlainEnsembleDynamics, self).__init__(
            transition=transition, learned_reward=learned_reward, reward_mech=reward_mech,
            learned_termination=learned_termination, termination_mech=termination_mech, optim_lr=optim_lr,
            weight_decay=weight_decay, optim_eps=optim_eps, logger=logger,
        )  # fmt: skip

    def _before_mech(self, mech: str):
        pass

    def _val_loss(self, val_dataset, mech):
        """[E] hold-out MSE = evaluate(...).mean(dim=(1,2)) (plain_dynamics.py:74,82), reduced on the device."""
        return self.evaluate_member_mse(val_dataset, mech=mech)

    def learn(
        self,
        replay_buffer,
        validation_ratio: float = 0.2,
        batch_size: int = 256,
        shuffle_each_epoch: bool = True,
        bootstrap_permutes: bool = False,
        longest_epoch: int = -1,
        improvement_threshold: float = 0.1,
        patience: int = 5,
        work_dir: Optional[Union[str, pathlib.Path]] = None,
        **kwargs
    ):
        """Early-stopped training of every learned mechanism, then elite selection by best hold-out MSE
        (plain_dynamics.py:44-111)."""
        train_dataset, val_dataset = self.dataset_split(replay_buffer, validation_ratio, batch_size,
                                                        shuffle_each_epoch, bootstrap_permutes)  # fmt: skip
        for mech in self.learn_mech:
            self._before_mech(mech)
            best_weights: Optional[Dict] = None
            epoch_iter = range(longest_epoch) if (longest_epoch is not None and longest_epoch > 0) else itertools.count()
            epochs_since_update = 0

            best_val_loss = self._val_loss(val_dataset, mech)

            for epoch in epoch_iter:
                train_loss = torch.tensor(self.train_epoch_device(train_dataset, mech=mech))
                val_loss = self._val_loss(val_dataset, mech)

                maybe_best_weights = self.maybe_get_best_weights(best_val_loss, val_loss, mech, improvement_threshold)
                if maybe_best_weights:
                    best_val_loss = torch.minimum(best_val_loss, val_loss)
                    best_weights = maybe_best_weights
                    epochs_since_update = 0
                else:
                    epochs_since_update += 1

                self.total_epoch[mech] += 1
                if self.logger is not None:
                    self.logger.record("{}/epoch".format(mech), epoch)
                    self.logger.record("{}/train_dataset_size".format(mech), train_dataset.num_stored)
                    self.logger.record("{}/val_dataset_size".format(mech), val_dataset.num_stored)
                    self.logger.record("{}/train_loss".format(mech), train_loss.mean().item())
                    self.logger.record("{}/val_loss".format(mech), val_loss.mean().item())
                    self.logger.record("{}/best_val_loss".format(mech), best_val_loss.mean().item())
                    self.logger.dump(self.total_epoch[mech])

                if patience and epochs_since_update >= patience:
                    break

            self.maybe_set_best_weights_and_elite(best_weights, best_val_loss, mech=mech)
        if work_dir is not None:
            self.save(work_dir)

    def maybe_get_best_weights(self, best_val_loss: torch.Tensor, val_loss: torch.Tensor, mech: str = "transition",
                               threshold: float = 0.01):  # fmt: skip
        """Snapshot of the WHOLE state_dict as soon as any member improved by more than `threshold`
        (plain_dynamics.py:113-127, quirk 7)."""
        improvement = (best_val_loss - val_loss) / torch.abs(best_val_loss)
        if (improvement > threshold).any().item():
            return copy.deepcopy(getattr(self, mech).state_dict())
        return None

    def maybe_set_best_weights_and_elite(self, best_weights: Optional[Dict], best_val_loss: torch.Tensor,
                                         mech: str = "transition"):  # fmt: skip
        model = getattr(self, mech)
        assert isinstance(model, EnsembleMLP)
        if best_weights is not None:
            model.load_state_dict(best_weights)
        sorted_indices = np.argsort(best_val_loss.tolist())
        model.set_elite_members(sorted_indices[: model.elite_num])


class ConstraintBasedDynamics(PlainEnsembleDynamics):
    """PlainEnsembleDynamics whose masked mechanisms get an oracle causal mask before training
    (constraint_based_dynamics.py:22-183; its CMI-test network is commented out in the reference)."""

    def __init__(
        self,
        transition: BaseTransition,
        learned_reward: bool = True,
        reward_mech: Optional[BaseRewardMech] = None,
        learned_termination: bool = False,
        termination_mech: Optional[BaseTerminationMech] = None,
        # trainer
        optim_lr: float = 1e-4,
        weight_decay: float = 1e-5,
        optim_eps: float = 1e-8,
        logger=None,
    ):
        super().__init__(
            transition=transition, learned_reward=learned_reward, reward_mech=reward_mech,
            learned_termination=learned_termination, termination_mech=termination_mech, optim_lr=optim_lr,
            weight_decay=weight_decay, optim_eps=optim_eps, logger=logger,
        )  # fmt: skip
        for mech in self.learn_mech:
            if hasattr(getattr(self, mech), "input_mask"):
                setattr(self, "{}_oracle_mask".format(mech), None)
                setattr(self, "{}_history_mask".format(mech),
                        torch.ones(getattr(self, mech).input_mask.shape).to(self.device))  # fmt: skip

    def build_cmi_test(self):
        """The conditional-mutual-information test network of the transition's shape (constraint_based_dynamics.py:66-78; the
        reference constructor has the call commented out, so nothing trains it by default)."""
        from .cmi_test import TransitionConditionalMutualInformationTest

        self.cmi_test = TransitionConditionalMutualInformationTest(
            obs_size=self.transition.obs_size, action_size=self.transition.action_size, ensemble_num=1, elite_num=1,
            residual=self.transition.residual, learn_logvar_bounds=self.transition.learn_logvar_bounds, num_layers=4,
            hid_size=200, activation_fn_cfg=self.transition.activation_fn_cfg, device=self.transition.device,
        )  # fmt: skip

    def set_oracle_mask(self, mech: str, mask):
        assert hasattr(self, "{}_oracle_mask".format(mech))
        setattr(self, "{}_oracle_mask".format(mech), to_tensor(mask))

    def _before_mech(self, mech: str):
        if hasattr(self, "{}_oracle_mask".format(mech)):
            getattr(self, mech).set_input_mask(getattr(self, "{}_oracle_mask".format(mech)))

    def learn(self, replay_buffer, validation_ratio: float = 0.2, batch_size: int = 256,
              shuffle_each_epoch: bool = True, bootstrap_permutes: bool = False, longest_epoch: Optional[int] = None,
              improvement_threshold: float = 0.1, patience: int = 5,
              work_dir: Optional[Union[str, pathlib.Path]] = None, **kwargs):  # fmt: skip
        super().learn(replay_buffer, validation_ratio, batch_size, shuffle_each_epoch, bootstrap_permutes,
                      longest_epoch, improvement_threshold, patience, None, **kwargs)  # fmt: skip
        self.save(work_dir)  # the reference saves unconditionally here (constraint_based_dynamics.py:152)
