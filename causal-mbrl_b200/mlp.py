"""EnsembleMLP: elite bookkeeping, losses and save/load with the reference's interface
(cmrl/models/networks/mlp.py:13-100), plus the machinery shared by every mechanism of this package:

* `_build_stack`   -- `hidden_layers` / `mean_and_logvar` modules with the reference's state_dict keys;
* `_MechFn`        -- one autograd node for the whole MLP + Gaussian head (K1 + K2 forward, K1/K2 backward);
* flat parameter arena -- all parameters are views of one contiguous fp32 buffer (and their grads of a
  second one), so Adam and the data-parallel all-reduce are single launches.
"""
import importlib
import pathlib
from typing import Dict, Optional, Sequence, Union

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .layers import EnsembleLinearLayer, ParallelEnsembleLinearLayer


def instantiate_activation(activation_fn_cfg):
    """`hydra.utils.instantiate(activation_fn_cfg)` (plain_transition.py:69-73) without requiring hydra:
    resolves `_target_` and calls it with the remaining keys.  ReLU when the config is None."""
    if activation_fn_cfg is None:
        return nn.ReLU()
    cfg = dict(activation_fn_cfg)
    target = cfg.pop("_target_")
    module, _, name = target.rpartition(".")
    return getattr(importlib.import_module(module), name)(**cfg)


def activation_kind(module: nn.Module) -> int:
    if isinstance(module, nn.ReLU):
        return ops.ACT_RELU
    if isinstance(module, nn.SiLU):
        return ops.ACT_SILU
    if isinstance(module, nn.Identity):
        return ops.ACT_IDENTITY
    raise NotImplementedError(
        "cmrl_b200 fuses the activation into the GEMM epilogue and implements ReLU, SiLU and Identity; got %r"
        % (module,)
    )


class _NllLossFn(torch.autograd.Function):
    """gaussian_nll(reduce=False) + reg (util.py:14-37, mlp.py:72-76) on the K3 kernels."""

    @staticmethod
    def forward(ctx, mean, logvar, target, reg):
        ctx.save_for_backward(mean, logvar, target)
        return ops.nll_loss_fwd(mean, logvar, target, reg)

    @staticmethod
    def backward(ctx, g_loss):
        mean, logvar, target = ctx.saved_tensors
        g_mean, g_logvar = ops.nll_loss_bwd(mean, logvar, target, g_loss=g_loss.contiguous())
        return g_mean, g_logvar, None, None


class _MseLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mean, target):
        ctx.save_for_backward(mean, target)
        return ops.mse_loss_fwd(mean, target)

    @staticmethod
    def backward(ctx, g_loss):
        mean, target = ctx.saved_tensors
        return ops.mse_loss_bwd(mean, target, g_loss), None


class _NoGrad:
    """marks a model whose forward was called under torch.no_grad()"""

    def __init__(self, model):
        self.model = model


class _MechFn(torch.autograd.Function):
    """(obs, action) -> (mean, logvar) for one mechanism: 4x(K1 linear + fused activation) -> K1 head
    GEMM -> K2 Gaussian head.  Backward replays the chain with the K1/K2 backward kernels."""

    @staticmethod
    def forward(ctx, model, obs, act, *params):
        ops._req(obs, "batch_obs"), ops._req(act, "batch_action")
        need_grad = any(ctx.needs_input_grad)
        if isinstance(model, _NoGrad):  # called under torch.no_grad(): inference, stays on the fp32 FFMA tier
            model, need_grad = model.model, False
        E, B = model.ensemble_num, obs.shape[-2]
        shared = obs.dim() == 2 or obs.shape[0] == 1 or (obs.stride(0) == 0 and act.stride(0) == 0)
        if obs.dim() == 4:  # one observation set per sub-net (CMI_test.py:103-105): [P,E,B,O]
            obs_c = obs.contiguous()
            x = torch.cat([obs_c, act.contiguous().expand(obs.shape[0], -1, -1, -1)], dim=-1).contiguous()
            res_obs = obs_c
            shared = False
        elif shared:  # same rows for every member (query / evaluate): never replicated
            o2 = obs if obs.dim() == 2 else obs[0]
            a2 = act if act.dim() == 2 else act[0]
            x = ops.gather_inputs(o2.contiguous(), a2.contiguous(), None)
            res_obs = o2
        else:
            obs_c, act_c = obs.contiguous(), act.contiguous()
            x = ops.gather_inputs(obs_c.view(E * B, -1), act_c.view(E * B, -1), None).view(E, B, -1)
            res_obs = obs_c
        mask = model._kernel_mask()
        mask_str = ops.mask_strides(mask, model._parallel_num(), E, B) if mask is not None else (0, 0, 0)
        n_layers = (len(params) - model._n_bound_params()) // 2 - 1
        kind = model._act_kind
        hs, zs = [x], []
        h = x
        for i in range(n_layers):
            out = ops.ens_linear_fwd(h, params[2 * i], params[2 * i + 1], kind, mask if i == 0 else None,
                                     mask_str, want_z=need_grad)  # fmt: skip
            h, z = out if need_grad else (out, None)
            hs.append(h)
            zs.append(z)
        raw = ops.ens_linear_fwd(h, params[2 * n_layers], params[2 * n_layers + 1], ops.ACT_IDENTITY)
        mean, logvar = model._head_fwd(raw, res_obs if model._residual() else None, E, B)
        if need_grad:
            ctx.model, ctx.n_layers, ctx.mask, ctx.mask_str, ctx.shared = model, n_layers, mask, mask_str, shared
            ctx.save_for_backward(raw, *hs, *zs, *params)
        return mean, logvar

    @staticmethod
    def backward(ctx, g_mean, g_logvar):
        model, L = ctx.model, ctx.n_layers
        saved = ctx.saved_tensors
        raw, hs, zs, params = saved[0], saved[1:L + 2], saved[L + 2:2 * L + 2], saved[2 * L + 2:]
        E, B = model.ensemble_num, raw.shape[-2]
        mn, mx = model._bounds()
        n_bounds = model._n_bound_params()
        grads = [None] * len(params)
        d_min = d_max = None
        if n_bounds and (ctx.needs_input_grad[3 + 2 * (L + 1)] or ctx.needs_input_grad[3 + 2 * (L + 1) + 1]):
            d_min, d_max = torch.zeros_like(mn), torch.zeros_like(mx)
            grads[2 * (L + 1)], grads[2 * (L + 1) + 1] = d_min, d_max
        dz = model._head_bwd(raw, E, B, g_mean, None if model.deterministic else g_logvar, d_min, d_max)
        kind = model._act_kind
        for i in range(L, -1, -1):
            w = params[2 * i]
            h_in = hs[i]
            if i == 0 and h_in.dim() < w.dim():  # shared / un-replicated input rows
                h_in = h_in.expand(tuple(w.shape[:-2]) + tuple(h_in.shape[-2:]))
            gw, gb = ops.ens_linear_bwd_dw(h_in, dz, w.shape, ctx.mask if i == 0 else None, ctx.mask_str)
            grads[2 * i], grads[2 * i + 1] = gw, gb
            if i > 0:
                dz = ops.ens_linear_bwd_dx(dz, w, zs[i - 1], kind)
        g_obs = g_act = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            # gradient w.r.t. the inputs (only multi-step training through ForwardEulerTransition needs it)
            dx = ops.ens_linear_bwd_dx(dz, params[0], None, ops.ACT_IDENTITY)
            if ctx.mask is not None:
                dx = (dx * model._broadcast_mask(E, B)).sum(dim=0)
            if ctx.shared:
                dx = dx.sum(dim=0, keepdim=True)
            O = model.obs_size
            g_obs, g_act = dx[..., :O], dx[..., O:]
            if model._residual() and g_mean is not None:
                g_obs = g_obs + (g_mean.sum(dim=0, keepdim=True) if ctx.shared else g_mean)
        return (None, g_obs, g_act) + tuple(grads)


class EnsembleMLP(nn.Module):
    _MODEL_FILENAME = "ensemble_mlp.pth"
    _head_layout = ops.HEAD_PLAIN

    def __init__(self, ensemble_num: int = 7, elite_num: int = 5, device: Union[str, torch.device] = "cpu"):
        super(EnsembleMLP, self).__init__()
        self.ensemble_num = ensemble_num
        self.elite_num = elite_num
        self.device = device

        self._elite_members: Optional[Sequence[int]] = np.random.permutation(ensemble_num)[:elite_num]
        self._model_save_attrs = ["elite_members", "state_dict"]
        self._flat = None  # flat parameter arena (device models only)
        self._flat_grad = None
        self._packed = {}  # private low-precision weight caches, keyed by kind -> (version, handle)
        self._weights_epoch = 0  # bumped by writers that bypass torch's version counters (the fused Adam kernel)

    # ------------------------------------------------------------------ elite bookkeeping (mlp.py:31-44)
    def set_elite_members(self, elite_indices: Sequence[int]):
        if len(elite_indices) != self.ensemble_num:
            assert len(elite_indices) == self.elite_num
            self._elite_members = list(elite_indices)

    @property
    def elite_members(self):
        return self._elite_members

    def get_random_index(self, batch_size: int, numpy_generator: Optional[np.random.Generator] = None):
        if numpy_generator:
            return numpy_generator.choice(self._elite_members, size=batch_size)
        return np.random.choice(self._elite_members, size=batch_size)

    # ------------------------------------------------------------------ save / load (mlp.py:46-63)
    def save(self, save_dir: Union[str, pathlib.Path]):
        """Saves {"elite_members", "state_dict", + registered attrs} to save_dir/_MODEL_FILENAME."""
        model_dict = {}
        for attr in self._model_save_attrs:
            model_dict[attr] = self.state_dict() if attr == "state_dict" else getattr(self, attr)
        torch.save(model_dict, pathlib.Path(save_dir) / self._MODEL_FILENAME)

    def load(self, load_dir: Union[str, pathlib.Path], load_device: Optional[str] = None):
        # elite_members is a numpy array / list, hence weights_only=False (same trust model as the reference)
        model_dict = torch.load(pathlib.Path(load_dir) / self._MODEL_FILENAME, map_location=load_device, weights_only=False)
        for attr in model_dict:
            if attr == "state_dict":
                self.load_state_dict(model_dict["state_dict"])
            else:
                getattr(self, "set_" + attr)(model_dict[attr])

    def add_save_attr(self, attr: str):
        assert hasattr(self, attr), "Class must has attribute {}".format(attr)
        assert attr not in self._model_save_attrs, "Attribute {} has been in model-save-list".format(attr)
        self._model_save_attrs.append(attr)

    @property
    def save_attr(self):
        return self._model_save_attrs

    @property
    def model_file_name(self):
        return self._MODEL_FILENAME

    # ------------------------------------------------------------------ construction helpers
    def create_linear_layer(self, l_in, l_out):
        return EnsembleLinearLayer(l_in, l_out, ensemble_num=self.ensemble_num)

    def _build_stack(self, in_size, hid_size, num_layers, head_out, activation_fn_cfg):
        """hidden_layers = Sequential(Sequential(linear, act) x num_layers); mean_and_logvar = linear
        (plain_transition.py:75-91): module creation order == the reference's, so torch's RNG stream and
        the state_dict keys match."""
        blocks = [nn.Sequential(self.create_linear_layer(in_size, hid_size), instantiate_activation(activation_fn_cfg))]
        for _ in range(num_layers - 1):
            blocks.append(nn.Sequential(self.create_linear_layer(hid_size, hid_size), instantiate_activation(activation_fn_cfg)))
        self.hidden_layers = nn.Sequential(*blocks)
        self.mean_and_logvar = self.create_linear_layer(hid_size, head_out)
        self._act_kind = activation_kind(blocks[0][1])

    def _finish_init(self):
        from .layers import truncated_normal_init

        self.apply(truncated_normal_init)
        self.to(self.device)
        self._flatten_parameters()

    # hooks the subclasses specialise ------------------------------------------------------------------
    def _parallel_num(self):
        return 1

    def _kernel_mask(self):
        return None

    def _residual(self):
        return False

    def _grad_dims(self) -> int:
        """Output dimensions `loss.mean()` averages over (net shards of a transition average over the full model's)."""
        return self._head_dim

    def target_columns(self, target):
        return target

    def _n_bound_params(self):
        return 0 if self.deterministic else 2

    def _bounds(self):
        if self.deterministic:
            return None, None
        return self.min_logvar, self.max_logvar

    def _head_fwd(self, raw, res_obs, E, B):
        """K2: raw head output -> (mean, logvar); specialised by models with another output layout."""
        mn, mx = self._bounds()
        return ops.gaussian_head_fwd(raw, self._head_layout, E, B, self._head_dim, res_obs, mn, mx,
                                     deterministic=self.deterministic)  # fmt: skip

    def _head_bwd(self, raw, E, B, g_mean, g_logvar, d_min, d_max):
        mn, mx = self._bounds()
        return ops.gaussian_head_bwd(raw, self._head_layout, E, B, self._head_dim, mn, mx, g_mean, g_logvar, d_min, d_max)

    def _stack_params(self):
        ps = []
        for block in self.hidden_layers:
            ps += [block[0].weight, block[0].bias]
        ps += [self.mean_and_logvar.weight, self.mean_and_logvar.bias]
        if not self.deterministic:
            ps += [self.min_logvar, self.max_logvar]
        return ps

    def _mech_forward(self, batch_obs, batch_action):
        # the outer grad mode is read here: inside Function.forward autograd is always off
        model = self if torch.is_grad_enabled() else _NoGrad(self)
        return _MechFn.apply(model, batch_obs, batch_action, *self._stack_params())

    # ------------------------------------------------------------------ losses (mlp.py:68-76)
    def get_mse_loss(self, model_in: Dict[(str, torch.Tensor)], target: torch.Tensor) -> torch.Tensor:
        pred_mean, pred_logvar = self.forward(**model_in)
        return _MseLossFn.apply(pred_mean, target)

    def get_nll_loss(self, model_in: Dict[(str, torch.Tensor)], target: torch.Tensor) -> torch.Tensor:
        pred_mean, pred_logvar = self.forward(**model_in)
        if self.max_logvar.requires_grad or self.min_logvar.requires_grad:
            nll_loss = _NllLossFn.apply(pred_mean, pred_logvar, target, 0.0)
            return nll_loss + 0.01 * (self.max_logvar.sum() - self.min_logvar.sum())
        return _NllLossFn.apply(pred_mean, pred_logvar, target, self._nll_reg())

    def _nll_reg(self) -> float:
        """0.01 * (sum(max_logvar) - sum(min_logvar)), cached per parameter version (it is a constant
        unless learn_logvar_bounds)."""
        key = (self.max_logvar._version, self.min_logvar._version, self.max_logvar.data_ptr())
        cached = getattr(self, "_reg_cache", None)
        if cached is None or cached[0] != key:
            reg = 0.01 * float((self.max_logvar.detach().sum() - self.min_logvar.detach().sum()).item())
            cached = (key, reg)
            self._reg_cache = cached
        return cached[1]

    # ------------------------------------------------------------------ private fp16 weight cache (tcgen05 tier)
    def weights_token(self) -> int:
        """A number that grows whenever the parameters may have changed: in-place torch operations (`load_state_dict`,
        `copy_`, ...) bump the tensors' `_version`, the fused Adam kernel writes through raw pointers and bumps
        `_weights_epoch` (optim.py).  Caches derived from the weights (packed fp16 copies, captured step graphs that
        bake their addresses) compare it before they are reused."""
        plist = self.__dict__.get("_token_params")
        if plist is None:  # the Parameter objects are fixed after construction (.to() / the arena only re-home their data)
            plist = self.__dict__["_token_params"] = list(self.parameters())
        token = getattr(self, "_weights_epoch", 0)
        for p in plist:
            token += p._version
        return token

    def packed_f16(self):
        """Padded fp16 copy of the weights in the layout cmrl_mlp_rollout_f16 stages by TMA; rebuilt whenever a
        parameter (or the input mask) changed.  Private: never appears in state_dict / checkpoints."""
        params = self._stack_params()
        n_lin = len(params) - self._n_bound_params()
        mask = self._kernel_mask()
        key = tuple(p._version for p in params[:n_lin]) + tuple(p.data_ptr() for p in params[:n_lin]) + (
            None if mask is None else (mask.data_ptr(), mask._version), getattr(self, "_weights_epoch", 0),
            bool(self.deterministic))  # fmt: skip
        cached = self._packed.get("f16")
        if cached is None or cached[0] != key:
            if getattr(self, "is_net_shard", False):
                raise ops._lib.CmrlB200Error("a net shard has no rollout kernels: gather the full model first (gather_from_shards)")
            E = self.ensemble_num
            if mask is not None and (mask.dim() == 4 or (mask.dim() == 3 and mask.shape[1] != E)):
                raise ops._lib.CmrlB200Error("per-sample input masks cannot be folded into packed weights; use precision='fp32'")
            mask_str = ops.mask_strides(mask, self._parallel_num(), E, -1) if mask is not None else (0, 0, 0)
            with torch.no_grad():
                inter = self._head_dim if (self._head_layout == ops.HEAD_PLAIN and not self.deterministic) else 0
                pack = ops.pack_mlp_f16(params[0:n_lin:2], params[1:n_lin:2], mask, mask_str, head_interleave_d=inter,
                                        act_kind=self._act_kind)
            cached = (key, pack)
            self._packed["f16"] = cached
        return cached[1]

    def packed_tf32(self):
        """3xTF32 operand blob for the fp32-grade tensor-core rollout (ops.Tf32RolloutPack; `.ok` False when the fused tf32
        kernel does not take this model); rebuilt whenever a parameter or the input mask changed."""
        params = self._stack_params()
        n_lin = len(params) - self._n_bound_params()
        mask = self._kernel_mask()
        key = tuple(p._version for p in params[:n_lin]) + tuple(p.data_ptr() for p in params[:n_lin]) + (
            None if mask is None else (mask.data_ptr(), mask._version), getattr(self, "_weights_epoch", 0),
            bool(self.deterministic))  # fmt: skip
        cached = self._packed.get("tf32")
        if cached is None or cached[0] != key:
            E = self.ensemble_num
            per_sample = mask is not None and (mask.dim() == 4 or (mask.dim() == 3 and mask.shape[1] != E))
            with torch.no_grad():
                pack = ops.Tf32RolloutPack(self) if not (per_sample or getattr(self, "is_net_shard", False)) else None
            cached = (key, pack)
            self._packed["tf32"] = cached
        return cached[1]

    # ------------------------------------------------------------------ flat parameter arena
    def _flatten_parameters(self):
        """Re-home every parameter as a view of one contiguous fp32 buffer (16-byte aligned slots).
        Trainable parameters come first so the optimiser touches flat[:_n_trainable] only."""
        params = list(self.parameters())
        if not params or not params[0].is_cuda:
            self._flat = self._flat_grad = None
            return
        order = [p for p in params if p.requires_grad] + [p for p in params if not p.requires_grad]
        slot, total, n_train = {}, 0, 0
        for p in order:
            slot[id(p)] = total
            total += (p.numel() + 3) // 4 * 4
            if p.requires_grad:
                n_train = total
        flat = torch.zeros(total, device=params[0].device, dtype=torch.float32)
        offsets = []
        for p in params:
            off = slot[id(p)]
            flat[off:off + p.numel()].view_as(p).copy_(p.data)
            p.data = flat[off:off + p.numel()].view_as(p)
            offsets.append(off)
        self._flat, self._flat_offsets, self._n_trainable = flat, offsets, n_train
        self._flat_grad = None

    def _arena_ok(self):
        if self._flat is None:
            return False
        base = self._flat.data_ptr()
        return all(p.data_ptr() == base + 4 * off for p, off in zip(self.parameters(), self._flat_offsets))

    def flat_parameters(self):
        """(flat_params, flat_grads): the arena; parameters' .grad are views of flat_grads."""
        if not self._arena_ok():
            self._flatten_parameters()
        if self._flat is None:
            raise ops._lib.CmrlB200Error("model parameters are not on a CUDA device")
        if self._flat_grad is None:
            self._flat_grad = torch.zeros_like(self._flat)
        for p, off in zip(self.parameters(), self._flat_offsets):
            if p.requires_grad:
                view = self._flat_grad[off:off + p.numel()].view_as(p)
                if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                    p.grad = view
        return self._flat, self._flat_grad


class ExternalMaskEnsembleMLP(EnsembleMLP):
    """Kept for interface parity with mlp.py:92-100 (an empty subclass in the reference)."""

    def __init__(self, ensemble_num: int = 7, elite_num: int = 5, device: Union[str, torch.device] = "cpu"):
        super().__init__(ensemble_num, elite_num, device)
