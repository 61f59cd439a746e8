"""torch-facing wrappers over the C ABI: pointer/stride extraction and output allocation only.

PyTorch is plumbing here (device memory, streams); every arithmetic op below is a kernel of
libcmrl_b200.so launched on torch's current CUDA stream.
"""
import contextlib
import gc

import torch

from . import _lib
from ._lib import ACT_IDENTITY, ACT_RELU, ACT_SILU, HEAD_PARALLEL, HEAD_PLAIN  # noqa: F401


@contextlib.contextmanager
def graph_capture(graph):
    """`torch.cuda.graph(graph)` with the cyclic garbage collector paused: a collection that fires in the middle of a
    capture can destroy an older CUDA graph / its memory pool (objects kept alive only by reference cycles), and the
    resulting cudaFree invalidates the capture in progress (cudaErrorStreamCaptureInvalidated)."""
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.graph(graph):
            yield
    finally:
        if was_enabled:
            gc.enable()


def _stream():
    # raw handle of torch's current stream on the current device (fast path: one C call)
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def _req(t, name, dtype=torch.float32):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.CmrlB200Error(
            "%s must be a CUDA tensor: cmrl_b200 runs on sm_100a kernels only, there is no CPU path" % name
        )
    if t.dtype != dtype:
        raise _lib.CmrlB200Error("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t


def _ptr(t):
    return None if t is None else t.data_ptr()


def _rows_contig(x, name):
    """x[..., B, C] with dense rows; leading dims may be broadcast (stride 0)."""
    if x.stride(-1) != 1 or (x.shape[-2] > 1 and x.stride(-2) != x.shape[-1]):
        x = x.contiguous()
    return x


def _lead_strides(x):
    s_e = x.stride(-3) if x.dim() >= 3 and x.shape[-3] > 1 else 0
    s_p = x.stride(-4) if x.dim() >= 4 and x.shape[-4] > 1 else 0
    return s_p, s_e


def mask_strides(mask, P, E, B):
    """Resolve the four mask shapes of ExternalMaskTransition.mask_input
    (external_mask_transition.py:116-139) to element strides (m_sp, m_se, m_sb)."""
    in_size = mask.shape[-1]
    if mask.dim() == 2:
        return in_size, 0, 0
    if mask.dim() == 3:
        if mask.shape[1] == E:  # [P,E,in]  (ensemble reading wins when E == B, reference quirk 5)
            return E * in_size, in_size, 0
        if mask.shape[1] == B:  # [P,B,in]
            return B * in_size, 0, in_size
        raise RuntimeError("input mask shape %a does not match x shape %a" % (tuple(mask.shape), (P, E, B, in_size)))
    if mask.dim() == 4:
        assert tuple(mask.shape[:3]) == (P, E, B)
        return E * B * in_size, B * in_size, in_size
    raise AssertionError("input mask must have 2..4 dims")


# GEMM tier of the TRAINING path (forward with saved pre-activations, dx, dw): 0 = fp32 FFMA, 1 = tcgen05 TF32 (1e-3
# tier), 3 = 3xTF32 error-compensated tcgen05 (fp32-grade), "auto" = 3 where the tensor-core kernel is faster than the
# FFMA kernel and 0 elsewhere (measured, profiles/tools/gemm_tier_time.py: forward and dx from ~1k rows per member on --
# 2.2x / 1.3-1.9x at 4k-64k rows -- and dw, whose K dimension is the batch, from 2k rows on with split-K, 3x; at the
# reference batch size of 256 rows both kernels are latency bound at ~17 us).  Inference forwards (query / evaluate / the fp32
# rollout tier) always run tier 0, whose row results do not depend on tile placement (selected-member rollouts stay
# bit-identical to the all-member path).
TRAIN_GEMM_TIER = "auto"
TRAIN_GEMM_TC_MIN_ROWS = 1024


def _train_tier(kind, rows, nets=1, k=0, n=0):
    """GEMM tier of a per-layer training call: 3 = tcgen05 3xTF32, 0 = fp32 FFMA.  Few nets: from ~1k rows per member (below
    that the launch is latency bound whichever pipe computes).  MANY nets (config C5: 1 600 nets of 512 x 512): the tensor-core
    tiles fill the chip at any batch size -- what counts is tiles = nets x row tiles x column tiles, and layers wide enough
    to fill a 128 x 256 tile."""
    if TRAIN_GEMM_TIER != "auto":
        return int(TRAIN_GEMM_TIER)
    wide = min(k, n) >= 128
    if kind == "dw":  # K of the dw GEMM is the batch: the tensor-core kernel pays once the batch is split over CTAs ...
        if rows >= 2048:
            return 3
        tiles = nets * -(-k // 128) * -(-n // 256)  # ... or there are enough (net, feature tile, column tile) items anyway
        return 3 if (wide and rows >= 256 and tiles >= 296) else 0
    if rows >= TRAIN_GEMM_TC_MIN_ROWS:
        return 3
    tiles = nets * -(-rows // 128) * -(-n // 256)
    return 3 if (wide and rows >= 128 and tiles >= 296) else 0


def ens_linear_fwd(x, w, b, act=ACT_IDENTITY, mask=None, mask_str=(0, 0, 0), want_z=False, row_offsets=None, tier=None):
    """y = act((x*mask) @ w + b).  w: [E,in,out] or [P,E,in,out]; x: [B,in] | [E,B,in] | [P,E,B,in]
    (broadcast dims allowed).  Ragged mode: x [N,in] flat + int32 row_offsets [E+1] -> y [N,out] / [P,N,out]."""
    _req(x, "x"), _req(w, "weight")
    parallel = w.dim() == 4
    P = w.shape[0] if parallel else 1
    E, in_size, out_size = w.shape[-3], w.shape[-2], w.shape[-1]
    if tier is None:
        tier = _train_tier("fwd", x.shape[-2], P * E, in_size, out_size) if (want_z and row_offsets is None) else 0
    w = w.contiguous()
    if b is not None:
        b = _req(b, "bias").contiguous()
    x = _rows_contig(x, "x")
    assert x.shape[-1] == in_size, "x last dim %d != in_size %d" % (x.shape[-1], in_size)
    B = x.shape[-2]
    if row_offsets is not None:
        assert x.dim() == 2 or (parallel and x.dim() == 3 and x.shape[0] == P)
        x_sp, x_se = (x.stride(0) if x.dim() == 3 else 0), 0
        out_shape = (P, B, out_size) if parallel else (B, out_size)
    else:
        x_sp, x_se = _lead_strides(x)
        out_shape = (P, E, B, out_size) if parallel else (E, B, out_size)
    y = torch.empty(out_shape, device=x.device, dtype=torch.float32)
    z = torch.empty_like(y) if want_z else None
    if mask is not None:
        mask = _req(mask, "mask").contiguous()
    _lib.call(
        "cmrl_ens_linear_fwd", x.data_ptr(), x_sp, x_se, w.data_ptr(), _ptr(b), y.data_ptr(), _ptr(z), P, E, B,
        in_size, out_size, _ptr(mask), mask_str[0], mask_str[1], mask_str[2], _ptr(row_offsets), act, tier, _stream(),
    )  # fmt: skip
    return (y, z) if want_z else y


def ens_linear_bwd_dx(dz, w, z_prev, act_prev, tier=None):
    """dx = (dz @ w^T) * act'(z_prev); dz [(P,)E,B,out] dense."""
    dz = _req(dz, "dz").contiguous()
    w = w.contiguous()
    parallel = w.dim() == 4
    P = w.shape[0] if parallel else 1
    E, in_size, out_size = w.shape[-3], w.shape[-2], w.shape[-1]
    tier = _train_tier("dx", dz.shape[-2], P * E, out_size, in_size) if tier is None else tier
    B = dz.shape[-2]
    dx = torch.empty(dz.shape[:-1] + (in_size,), device=dz.device, dtype=torch.float32)
    if z_prev is not None:
        z_prev = z_prev.contiguous()
    _lib.call(
        "cmrl_ens_linear_bwd_dx", dz.data_ptr(), w.data_ptr(), _ptr(z_prev), dx.data_ptr(), P, E, B, in_size,
        out_size, act_prev, tier, _stream(),
    )  # fmt: skip
    return dx


_SPLITK_WS = {}
_SPLITK_WS_RETIRED = []


def splitk_workspace(device, floats):
    """Persistent per-device scratch for the split-K dw GEMMs (they run one after the other on one stream).  It is grown
    outside CUDA-graph captures only: inside a capture a too small workspace means "no split" for that call (call
    `splitk_workspace(device, n)` before capturing, as the device-resident epoch runner does)."""
    device = torch.device(device)
    cur = _SPLITK_WS.get(device)
    if cur is not None and cur.numel() >= floats:
        return cur
    if torch.cuda.is_current_stream_capturing():
        return cur
    if cur is not None:
        _SPLITK_WS_RETIRED.append(cur)  # a captured step graph may hold its address: never freed while the process lives
    cur = torch.empty(floats, device=device, dtype=torch.float32)
    _SPLITK_WS[device] = cur
    return cur


def ens_linear_bwd_dw(x, dz, w_shape, mask=None, mask_str=(0, 0, 0), dw=None, db=None, want_db=True, accumulate=False,
                      tier=None):
    """dw = (x*mask)^T @ dz, db = sum_b dz, written into dw/db when given (flat-arena views)."""
    dz = _req(dz, "dz").contiguous()
    x = _rows_contig(_req(x, "x"), "x")
    parallel = len(w_shape) == 4
    P = w_shape[0] if parallel else 1
    E, in_size, out_size = w_shape[-3], w_shape[-2], w_shape[-1]
    tier = _train_tier("dw", dz.shape[-2], P * E, in_size, out_size) if tier is None else tier
    B = dz.shape[-2]
    x_sp, x_se = _lead_strides(x)
    if dw is None:
        dw = torch.empty(tuple(w_shape), device=dz.device, dtype=torch.float32)
        accumulate = False
    if db is None and want_db:
        db = torch.empty(tuple(w_shape[:-2]) + (1, out_size), device=dz.device, dtype=torch.float32)
    assert dw.is_contiguous() and (db is None or db.is_contiguous())
    ws, ws_floats = None, 0
    if tier and B >= 2048:  # split-K workspace: one partial [P*E, in+1, out] per batch slice (capped at 256 MB)
        per_slice = P * E * (in_size + 1) * out_size
        slices = min(32, (256 << 20) // (4 * per_slice))
        if slices >= 2:
            ws = splitk_workspace(dz.device, slices * per_slice)
            ws_floats = 0 if ws is None else ws.numel()
    _lib.call(
        "cmrl_ens_linear_bwd_dw", x.data_ptr(), x_sp, x_se, _ptr(mask), mask_str[0], mask_str[1], mask_str[2],
        dz.data_ptr(), dw.data_ptr(), _ptr(db), P, E, B, in_size, out_size, 1 if accumulate else 0, tier, _ptr(ws),
        ws_floats, _stream(),
    )  # fmt: skip
    return dw, db


def gaussian_head_fwd(raw, layout, E, B, D, obs, min_logvar, max_logvar, deterministic=False):
    raw = _req(raw, "raw").contiguous()
    mean = torch.empty((E, B, D), device=raw.device, dtype=torch.float32)
    logvar = None if deterministic else torch.empty_like(mean)
    obs_se, obs_ld = 0, 0
    if obs is not None:
        obs = _rows_contig(_req(obs, "obs"), "obs")
        obs_ld = obs.shape[-1]
        obs_se = _lead_strides(obs)[1]
    bounds_n = 0 if deterministic else min_logvar.numel()
    _lib.call(
        "cmrl_gaussian_head_fwd", raw.data_ptr(), layout, E, B, D, _ptr(obs), obs_se, obs_ld,
        None if deterministic else min_logvar.data_ptr(), None if deterministic else max_logvar.data_ptr(), bounds_n,
        mean.data_ptr(), _ptr(logvar), _stream(),
    )  # fmt: skip
    return mean, logvar


def gaussian_head_bwd(raw, layout, E, B, D, min_logvar, max_logvar, g_mean, g_logvar, d_min=None, d_max=None):
    raw = raw.contiguous()
    d_raw = torch.empty_like(raw)
    if g_mean is not None:
        g_mean = g_mean.contiguous()
    if g_logvar is not None:
        g_logvar = g_logvar.contiguous()
    bounds_n = 0 if min_logvar is None else min_logvar.numel()
    _lib.call(
        "cmrl_gaussian_head_bwd", raw.data_ptr(), layout, E, B, D, _ptr(min_logvar), _ptr(max_logvar), bounds_n,
        _ptr(g_mean), _ptr(g_logvar), d_raw.data_ptr(), _ptr(d_min), _ptr(d_max), _stream(),
    )  # fmt: skip
    return d_raw


def _target_stride(target, E):
    """target [E,B,D] (or broadcast over members) -> (contiguous-rows tensor, member stride)."""
    target = _req(target, "target")
    if target.dim() == 3 and target.shape[0] == E and target.stride(0) != 0 and E > 1:
        target = target.contiguous()
        return target, target.stride(0)
    if target.dim() == 3:  # broadcast over members
        t0 = target[0].contiguous()
        return t0, 0
    return target.contiguous(), 0


def nll_loss_fwd(mean, logvar, target, reg):
    mean, logvar = mean.contiguous(), logvar.contiguous()
    E = mean.shape[0]
    n = mean[0].numel()
    target, t_se = _target_stride(target, E)
    loss = torch.empty_like(mean)
    _lib.call("cmrl_nll_loss_fwd", mean.data_ptr(), logvar.data_ptr(), target.data_ptr(), t_se, E, n, float(reg),
              loss.data_ptr(), _stream())  # fmt: skip
    return loss


def nll_loss_bwd(mean, logvar, target, g_loss=None, g_scalar=0.0):
    mean, logvar = mean.contiguous(), logvar.contiguous()
    E = mean.shape[0]
    n = mean[0].numel()
    target, t_se = _target_stride(target, E)
    g_mean, g_logvar = torch.empty_like(mean), torch.empty_like(mean)
    if g_loss is not None:
        g_loss = g_loss.contiguous()
    _lib.call("cmrl_nll_loss_bwd", mean.data_ptr(), logvar.data_ptr(), target.data_ptr(), t_se, E, n, _ptr(g_loss),
              float(g_scalar), g_mean.data_ptr(), g_logvar.data_ptr(), _stream())  # fmt: skip
    return g_mean, g_logvar


def mse_loss_fwd(mean, target):
    mean = mean.contiguous()
    E = mean.shape[0]
    n = mean[0].numel()
    target, t_se = _target_stride(target, E)
    loss = torch.empty_like(mean)
    _lib.call("cmrl_mse_loss_fwd", mean.data_ptr(), target.data_ptr(), t_se, E, n, loss.data_ptr(), _stream())
    return loss


def mse_loss_bwd(mean, target, g_loss):
    mean, g_loss = mean.contiguous(), g_loss.contiguous()
    E = mean.shape[0]
    n = mean[0].numel()
    target, t_se = _target_stride(target, E)
    g_mean = torch.empty_like(mean)
    _lib.call("cmrl_mse_loss_bwd", mean.data_ptr(), target.data_ptr(), t_se, E, n, g_loss.data_ptr(), g_mean.data_ptr(), _stream())
    return g_mean


def mse_member_sums(mean, target, sums):
    """sums (float64 [E], caller-zeroed) += per-member sum of squared errors."""
    mean = mean.contiguous()
    E = mean.shape[0]
    n = mean[0].numel()
    target, t_se = _target_stride(target, E)
    assert sums.dtype == torch.float64 and sums.numel() == E
    _lib.call("cmrl_mse_member_sums", mean.data_ptr(), target.data_ptr(), t_se, E, n, sums.data_ptr(), _stream())
    return sums


def adam_l2_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    n = param.numel()
    assert param.is_contiguous() and grad.is_contiguous() and grad.numel() == n
    _lib.call("cmrl_adam_l2_step", param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), n,
              float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), float(grad_scale), int(step),
              _stream())  # fmt: skip


def adam_l2_step_dev(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step_dev, grad_scale=1.0):
    _lib.call("cmrl_adam_l2_step_dev", param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
              param.numel(), float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), float(grad_scale),
              step_dev.data_ptr(), _stream())  # fmt: skip


def sum_acc_f64(x, acc, workspace):
    """acc (float64 scalar tensor) += x.sum() on the device, deterministic; workspace: 66 zero-initialised doubles."""
    _lib.call("cmrl_sum_acc_f64", x.data_ptr(), x.numel(), acc.data_ptr(), workspace.data_ptr(), _stream())


def bootstrap_gather(obs, act, target, member_idx, order, start_dev, advance, B, out_obs, out_act, out_target):
    E, N = member_idx.shape
    _lib.call("cmrl_bootstrap_gather", obs.data_ptr(), act.data_ptr(), target.data_ptr(), member_idx.data_ptr(),
              order.data_ptr(), start_dev.data_ptr(), int(advance), E, B, N, obs.shape[1], act.shape[1], target.shape[1],
              out_obs.data_ptr(), out_act.data_ptr(), out_target.data_ptr(), _stream())  # fmt: skip


# ----------------------------------------------------------------------------- rollout
def draw_member_noise(seed, stream_id, N, elite_u8, D, want_noise=True, device=None, stream_base=None, out=None):
    """out: optional (idx [N] uint8, noise [N,D] float32 or None) to draw into"""
    device = elite_u8.device if device is None else device
    if out is not None:
        idx, noise = out
        assert idx.numel() == N and (noise is None) == (not want_noise)
    else:
        idx = torch.empty(N, device=device, dtype=torch.uint8)
        noise = torch.empty((N, D), device=device, dtype=torch.float32) if want_noise else None
    _lib.call("cmrl_draw_member_noise", int(seed), int(stream_id), _ptr(stream_base), N, elite_u8.data_ptr(),
              elite_u8.numel(), D, idx.data_ptr(), _ptr(noise), _stream())  # fmt: skip
    return idx, noise


def counter_add(counter, delta):
    _lib.call("cmrl_counter_add", counter.data_ptr(), int(delta), _stream())


def draw_uniform_index(seed, stream_id, N, upper, device, stream_base=None):
    out = torch.empty(N, device=device, dtype=torch.int32)
    _lib.call("cmrl_draw_uniform_index", int(seed), int(stream_id), _ptr(stream_base), N, int(upper), out.data_ptr(),
              _stream())  # fmt: skip
    return out


BUCKET_SCRATCH_WARPS = 512  # CMRL_BUCKET_SCRATCH_WARPS (include/cmrl_b200.h)


def member_bucket(member_idx, E, perm=None, offsets=None, counters=None):
    N = member_idx.numel()
    dev = member_idx.device
    perm = torch.empty(N, device=dev, dtype=torch.int32) if perm is None else perm
    offsets = torch.empty(E + 1, device=dev, dtype=torch.int32) if offsets is None else offsets
    counters = torch.empty(BUCKET_SCRATCH_WARPS * E, device=dev, dtype=torch.int32) if counters is None else counters
    assert counters.numel() >= BUCKET_SCRATCH_WARPS * E
    _lib.call("cmrl_member_bucket", member_idx.data_ptr(), N, E, perm.data_ptr(), offsets.data_ptr(),
              counters.data_ptr(), _stream())  # fmt: skip
    return perm, offsets


def gather_inputs(obs, act, perm):
    N, O = obs.shape
    A = act.shape[1]
    x = torch.empty((N, O + A), device=obs.device, dtype=torch.float32)
    _lib.call("cmrl_gather_inputs", obs.data_ptr(), act.data_ptr(), _ptr(perm), N, O, A, x.data_ptr(), _stream())
    return x


def gather_rows(x, rows):
    """x[rows] for a dense [N,C] float32 tensor and int32 row numbers on the device (cmrl_gather_inputs without an action part)."""
    assert x.dim() == 2 and x.is_contiguous() and rows.dtype == torch.int32
    out = torch.empty((rows.numel(), x.shape[1]), device=x.device, dtype=torch.float32)
    _lib.call("cmrl_gather_inputs", x.data_ptr(), None, rows.data_ptr(), rows.numel(), x.shape[1], 0, out.data_ptr(), _stream())
    return out


def sample_scatter(raw_sorted, layout, perm, N, D, obs, min_logvar, max_logvar, noise, out=None):
    pred = torch.empty((N, D), device=raw_sorted.device, dtype=torch.float32) if out is None else out
    bounds_n = 0 if min_logvar is None else min_logvar.numel()
    _lib.call("cmrl_sample_scatter", raw_sorted.data_ptr(), layout, _ptr(perm), N, D, _ptr(obs),
              0 if obs is None else obs.shape[-1], _ptr(min_logvar), _ptr(max_logvar), bounds_n, _ptr(noise),
              pred.data_ptr(), _stream())  # fmt: skip
    return pred


def select_sample(mean, logvar, member_idx, noise):
    E, N, D = mean.shape
    pred = torch.empty((N, D), device=mean.device, dtype=torch.float32)
    _lib.call("cmrl_select_sample", mean.data_ptr(), _ptr(logvar), member_idx.data_ptr(), E, N, D, _ptr(noise),
              pred.data_ptr(), _stream())  # fmt: skip
    return pred


def mopo_penalty(mean):
    E, N, O = mean.shape
    pen = torch.empty(N, device=mean.device, dtype=torch.float32)
    _lib.call("cmrl_mopo_penalty", mean.contiguous().data_ptr(), E, N, O, pen.data_ptr(), _stream())
    return pen


def env_step_tail(next_obs, reward, terminal_f, terminal_u8, penalty, penalty_coeff, env_len, max_episode_steps,
                  reset_obs, reset_index, obs_out, reward_out, done, truncated, terminal_obs, n_done):  # fmt: skip
    N, O = next_obs.shape
    _lib.call("cmrl_env_step_tail", next_obs.data_ptr(), reward.data_ptr(), _ptr(terminal_f), _ptr(terminal_u8),
              _ptr(penalty), float(penalty_coeff), env_len.data_ptr(), int(max_episode_steps), _ptr(reset_obs),
              _ptr(reset_index), N, O, obs_out.data_ptr(), reward_out.data_ptr(), done.data_ptr(),
              truncated.data_ptr(), _ptr(terminal_obs), _ptr(n_done), _stream())  # fmt: skip


# ----------------------------------------------------------------------------- tensor-core tier (tcgen05, fp16)
def _round_up(x, m):
    return (x + m - 1) // m * m


ACT_SILU_HALF = 3  # CMRL_ACT_SILU_HALF: SiLU on halved pre-activations (hidden layers packed with scale 0.5)


def pack_mlp_f16(weights, biases, mask=None, mask_str=(0, 0, 0), head_interleave_d=0, act_kind=None):
    """Pack a mechanism's fp32 parameters into the private fp16 device layout of cmrl_mlp_rollout_f16.
    weights/biases: per layer [(P,)E,in,out] / [(P,)E,1,out] (hidden layers, then the head).
    Returns (blob half tensor [nets, blob_halfs], layer_off, layer_k, layer_n) -- never part of state_dict."""
    w0 = weights[0]
    parallel = w0.dim() == 4
    P = w0.shape[0] if parallel else 1
    E = w0.shape[-3]
    layer_k, layer_n, layer_off, off = [], [], [], 0
    for w in weights:
        k_pad, n_pad = _round_up(w.shape[-2] + 1, 16), _round_up(w.shape[-1], 16)
        layer_k.append(k_pad)
        layer_n.append(n_pad)
        layer_off.append(off)
        off += k_pad * n_pad
    # Wide models (hidden a multiple of 16 whose padded width no longer fits the weight-stationary kernels) keep N = hidden:
    # the streamed-weights kernel (variant 6) carries the bias input in a K step of its own.  Otherwise hidden outputs
    # feed the next layer's K including the bias slot: widen N of hidden layers to the next K.
    hid = weights[0].shape[-1]
    wide = len(weights) > 1 and hid % 16 == 0 and hid > 207 and all(w.shape[-1] == hid for w in weights[:-1])
    if not wide:
        for i in range(len(weights) - 1):
            layer_n[i] = layer_k[i + 1]
    layer_off, off = [], 0
    for k_pad, n_pad in zip(layer_k, layer_n):
        layer_off.append(off)
        off += k_pad * n_pad
    blob = torch.zeros((P * E, off), device=w0.device, dtype=torch.float16)
    for i, (w, b) in enumerate(zip(weights, biases)):
        w, b = w.detach().contiguous(), b.detach().contiguous()
        m = mask if i == 0 else None
        inter = head_interleave_d if i == len(weights) - 1 else 0
        # SiLU models: hidden layers are stored halved (exact), the kernels evaluate h + h*tanh(h) on h = z/2
        scale = 0.5 if (act_kind == ACT_SILU and i < len(weights) - 1) else 1.0
        _lib.call("cmrl_pack_layer_f16", w.data_ptr(), b.data_ptr(), _ptr(m), mask_str[0], mask_str[1], P, E,
                  w.shape[-2], w.shape[-1], layer_k[i], layer_n[i], inter, scale, blob.data_ptr() + 2 * layer_off[i], off,
                  _stream())  # fmt: skip
    # [blob, layer offsets, K, N, ctypes views (lazy), unused, hidden layers halved?]
    return [blob, layer_off, layer_k, layer_n, None, None, act_kind == ACT_SILU]


# Kernel variant of the tensor-core rollout tier.  None (default) = chosen by the library from the model's shape: 5 (CTA
# pairs, two 256-row tiles in flight, A operand partly in tensor memory) where two tiles' accumulators and heads fit tensor
# memory, 3 (pairs, two 128-row tiles in flight) for wider heads, 6 (pairs, weights streamed in K chunks) for hidden layers
# that do not fit shared memory (a multiple of 16 from 208 to 512).  Other widths use precision="fp32".  Tests pin a variant.
TC_VARIANT = None


def mlp_rollout_f16(pack, P, E, obs, act, perm, offsets, H, head_dim, layout, gaussian, act_kind, residual,
                    min_logvar, max_logvar, noise, out=None, all_members=False):  # fmt: skip
    import ctypes

    blob, layer_off, layer_k, layer_n = pack[:4]
    N, O = obs.shape
    A = act.shape[1]
    L = len(layer_k) - 1
    if pack[4] is None:  # ctypes views of the layer tables, built once per pack
        arr = ctypes.c_int * (L + 1)
        pack[4] = (arr(*layer_off), arr(*layer_k), arr(*layer_n))
    c_off, c_k, c_n = pack[4]
    if pack[6]:
        assert act_kind == ACT_SILU, "this pack stores halved hidden layers: it belongs to a SiLU model"
        act_kind = ACT_SILU_HALF
    # all_members: perm / offsets are the all-member buckets (E * N rows), the output is [E,N,D] (MOPO penalty: every member's mean)
    rows = perm.numel()
    if out is None:
        out = torch.empty(((E, N, head_dim) if all_members else (N, head_dim)), device=obs.device, dtype=torch.float32)
    pred = out
    bounds_n = 0 if min_logvar is None else min_logvar.numel()
    _lib.call("cmrl_mlp_rollout_f16", blob.data_ptr(), blob.shape[1], c_off, c_k, c_n, L,
              P, E, rows, O, A, H, head_dim, layout, 1 if gaussian else 0, act_kind, obs.data_ptr(), act.data_ptr(),
              perm.data_ptr(), offsets.data_ptr(), 1 if residual else 0, _ptr(min_logvar), _ptr(max_logvar),
              bounds_n, _ptr(noise), pred.data_ptr(), N * head_dim if all_members else 0,
              0 if TC_VARIANT is None else int(TC_VARIANT), _stream())  # fmt: skip
    return pred


def mlp_rollout_f16_groups(groups):
    """ONE launch of the fused f16 rollout kernel for several mechanisms of a step.  groups: list of dicts with the arguments of
    `mlp_rollout_f16` (pack, P, E, obs, act, perm, offsets, H, head_dim, layout, gaussian, act_kind, residual, min_logvar,
    max_logvar, noise, out).  Returns the outputs, or None when the mechanisms cannot share a launch (different packed layer
    tables, a shape the pair kernel does not take): the caller then launches them one by one."""
    import ctypes

    first = groups[0]["pack"]
    arr = (_lib.RolloutGroup * len(groups))()
    keep, outs = [], []
    for i, gd in enumerate(groups):
        pack = gd["pack"]
        blob, layer_off, layer_k, layer_n = pack[:4]
        if (list(layer_off), list(layer_k), list(layer_n), bool(pack[6])) != (list(first[1]), list(first[2]), list(first[3]), bool(first[6])):
            return None
        L = len(layer_k) - 1
        if pack[4] is None:
            c_arr = ctypes.c_int * (L + 1)
            pack[4] = (c_arr(*layer_off), c_arr(*layer_k), c_arr(*layer_n))
        c_off, c_k, c_n = pack[4]
        obs, act = gd["obs"], gd["act"]
        N, O = obs.shape
        act_kind = gd["act_kind"]
        if pack[6]:
            assert act_kind == ACT_SILU
            act_kind = ACT_SILU_HALF
        out = gd.get("out")
        if out is None:
            out = torch.empty((N, gd["head_dim"]), device=obs.device, dtype=torch.float32)
        mn, mx, noise = gd["min_logvar"], gd["max_logvar"], gd["noise"]
        g = arr[i]
        g.packed, g.blob_halfs = blob.data_ptr(), blob.shape[1]
        g.layer_off_halfs, g.layer_k, g.layer_n = (ctypes.cast(c_off, ctypes.c_void_p), ctypes.cast(c_k, ctypes.c_void_p),
                                                   ctypes.cast(c_n, ctypes.c_void_p))
        g.num_layers, g.P, g.E, g.N, g.O, g.A, g.H = L, gd["P"], gd["E"], gd["perm"].numel(), O, act.shape[1], gd["H"]
        g.head_dim, g.layout, g.gaussian, g.act = gd["head_dim"], gd["layout"], 1 if gd["gaussian"] else 0, act_kind
        g.residual, g.bounds_n = 1 if gd["residual"] else 0, 0 if mn is None else mn.numel()
        g.obs, g.actn, g.perm, g.offsets = obs.data_ptr(), act.data_ptr(), gd["perm"].data_ptr(), gd["offsets"].data_ptr()
        g.min_logvar, g.max_logvar, g.noise, g.pred = _ptr(mn), _ptr(mx), _ptr(noise), out.data_ptr()
        keep.append((obs, act, mn, mx, noise))
        outs.append(out)
    lib = _lib.load()
    rc = lib.cmrl_mlp_rollout_f16_groups(ctypes.byref(arr), len(groups), _stream())
    if rc == -2:
        return None
    if rc != 0:
        raise _lib.CmrlB200Error("cmrl_mlp_rollout_f16_groups failed (code %d): %s" % (rc, lib.cmrl_last_error().decode("utf-8", "replace")))
    return outs


class Tf32RolloutPack:
    """3xTF32 operand blob of a model for cmrl_mlp_rollout_tf32 (the layout cmrl_train_pack_tf32 writes), or ok == False
    when the fused tf32 kernel does not take the model (deterministic head, wide hidden layers, more than 16 head outputs)."""

    def __init__(self, model):
        import ctypes

        params = model._stack_params()
        n_lin = len(params) - model._n_bound_params()
        ws, bs = params[0:n_lin:2], params[1:n_lin:2]
        self.L = len(ws) - 1
        self.P, self.E = model._parallel_num(), model.ensemble_num
        self.in_size, self.hidden = ws[0].shape[-2], ws[0].shape[-1]
        self.layout, self.head_dim = model._head_layout, model._head_dim
        head_outputs = 1 if self.layout == HEAD_PARALLEL else self.head_dim
        blob = ctypes.c_longlong(0)
        self.ok = (not model.deterministic and self.E <= 64
                   and ((self.layout == HEAD_PLAIN and self.P == 1) or (self.layout == HEAD_PARALLEL and self.P == self.head_dim))
                   and _lib.load().cmrl_train_fused_sizes(self.L, self.in_size, self.hidden, head_outputs, self.P, self.E, 1,
                                                          ctypes.byref(blob), None, None) == 0)
        if not self.ok:
            return
        self.packed = torch.empty(self.P * self.E * blob.value, device=ws[0].device, dtype=torch.float32)
        arr = ctypes.c_void_p * (self.L + 1)
        mask = model._kernel_mask()
        m_str = mask_strides(mask, self.P, self.E, -1)[:2] if mask is not None else (0, 0)
        _lib.call("cmrl_train_pack_tf32", arr(*[w.data_ptr() for w in ws]), arr(*[b.data_ptr() for b in bs]), self.L, self.P, self.E,
                  self.in_size, self.hidden, head_outputs, _ptr(mask), m_str[0], m_str[1], self.packed.data_ptr(), _stream())


def mlp_rollout_tf32(pack, obs, act, perm, offsets, act_kind, residual, min_logvar, max_logvar, noise, out=None, all_members=False):
    """Fused selected-member rollout on tcgen05 kind::tf32 with hi/lo split operands (fp32-grade): see cmrl_mlp_rollout_tf32.
    all_members: perm / offsets are the all-member buckets (E * N rows) and the output is [E,N,D]."""
    N, O = obs.shape
    if out is None:
        out = torch.empty(((pack.E, N, pack.head_dim) if all_members else (N, pack.head_dim)), device=obs.device, dtype=torch.float32)
    _lib.call("cmrl_mlp_rollout_tf32", pack.packed.data_ptr(), pack.L, pack.P, pack.E, perm.numel(), O, act.shape[1], pack.hidden,
              pack.head_dim, pack.layout, act_kind, 1 if residual else 0, obs.data_ptr(), act.data_ptr(), perm.data_ptr(),
              offsets.data_ptr(), _ptr(min_logvar), _ptr(max_logvar), 0 if min_logvar is None else min_logvar.numel(), _ptr(noise),
              out.data_ptr(), N * pack.head_dim if all_members else 0, _stream())
    return out


# ----------------------------------------------------------------------------- fused tcgen05 training step
# One optimiser step of a mechanism = pack + fused forward / loss / dz chain + all weight gradients (three launches of
# train_tc.cu) instead of ~25 per-layer launches.  TRAIN_FUSED_PASSES: 3 = 3xTF32 (fp32-grade, the default), 1 = TF32.
TRAIN_FUSED = True
TRAIN_FUSED_PASSES = 3
TRAIN_FUSED_SCRATCH_CAP = 24 << 30  # bytes of saved activations above which the per-layer path is used instead
TRAIN_FUSED_DW_WS_CAP = 256 << 20   # bytes of batch-slice partial sums for the weight-gradient GEMMs


class FusedTrainPlan:
    """Device buffers and host pointer tables of the fused training step for one model and batch size."""

    def __init__(self, model, E, B):
        import ctypes

        params = model._stack_params()
        n_lin = len(params) - model._n_bound_params()
        ws, bs = params[0:n_lin:2], params[1:n_lin:2]
        self.L = len(ws) - 1
        self.P, self.E, self.B = model._parallel_num(), E, B
        self.in_size, self.hidden = ws[0].shape[-2], ws[0].shape[-1]
        self.layout = model._head_layout
        self.head_dim = model._head_dim
        self.head_outputs = 1 if self.layout == HEAD_PARALLEL else self.head_dim
        self.act = model._act_kind
        self.residual = 1 if model._residual() else 0
        blob, scratch, dw_ws = ctypes.c_longlong(0), ctypes.c_longlong(0), ctypes.c_longlong(0)
        rc = _lib.load().cmrl_train_fused_sizes(self.L, self.in_size, self.hidden, self.head_outputs, self.P, E, B,
                                                 ctypes.byref(blob), ctypes.byref(scratch), ctypes.byref(dw_ws))
        self.ok = rc == 0 and scratch.value * 4 <= TRAIN_FUSED_SCRATCH_CAP
        if not self.ok:
            return
        dev = ws[0].device
        nets = self.P * E
        self.packed = torch.empty(nets * blob.value, device=dev, dtype=torch.float32)
        self.scratch = torch.empty(scratch.value, device=dev, dtype=torch.float32)
        tiles = (B + 127) // 128
        slices = min(64, tiles, TRAIN_FUSED_DW_WS_CAP // (4 * dw_ws.value)) if tiles > 1 else 0
        self.dw_ws = torch.empty(slices * dw_ws.value, device=dev, dtype=torch.float32) if slices >= 2 else None
        arr = ctypes.c_void_p * (self.L + 1)
        self.w_ptrs = arr(*[w.data_ptr() for w in ws])
        self.b_ptrs = arr(*[b.data_ptr() for b in bs])
        self.dw_ptrs = arr(*[w.grad.data_ptr() for w in ws])
        self.db_ptrs = arr(*[b.grad.data_ptr() for b in bs])
        mask = model._kernel_mask()
        self.mask = mask
        self.mask_str = mask_strides(mask, self.P, E, -1)[:2] if mask is not None else (0, 0)
        self.launches = 3 + (1 if self.dw_ws is not None else 0)

    def run_pack(self):
        m = self.mask
        _lib.call("cmrl_train_pack_tf32", self.w_ptrs, self.b_ptrs, self.L, self.P, self.E, self.in_size, self.hidden,
                  self.head_outputs, _ptr(m), self.mask_str[0], self.mask_str[1], self.packed.data_ptr(), _stream())  # fmt: skip

    def run_fused(self, obs, act, target, mn, mx, reg, g_scale, loss=None):
        E, B = self.E, self.B
        _lib.call("cmrl_train_fused", self.packed.data_ptr(), self.L, self.P, E, B, obs.shape[-1], act.shape[-1], self.hidden,
                  self.head_dim, self.layout, self.act, self.residual, int(TRAIN_FUSED_PASSES), obs.data_ptr(),
                  _lead_strides(obs)[1], act.data_ptr(), _lead_strides(act)[1], target.data_ptr(), _lead_strides(target)[1],
                  target.shape[-1], mn.data_ptr(), mx.data_ptr(), mn.numel(), float(reg), float(g_scale), _ptr(loss),
                  self.scratch.data_ptr(), _stream())  # fmt: skip

    def run_dw(self, *unused):
        m = self.mask
        _lib.call("cmrl_train_dw", self.scratch.data_ptr(), self.L, self.P, self.E, self.B, self.in_size, self.hidden,
                  self.head_outputs, self.act, self.dw_ptrs, self.db_ptrs, _ptr(m), self.mask_str[0], self.mask_str[1],
                  _ptr(self.dw_ws), 0 if self.dw_ws is None else self.dw_ws.numel(), _stream())  # fmt: skip

    def step(self, obs, act, target, mn, mx, reg, g_scale, want_loss=True, after_pack=None, after_fused=None):
        """pack -> fused forward / NLL / dz chain -> weight gradients, all on the current stream; returns the loss.
        after_pack: called between the packing (which reads only the weights) and the first kernel that reads the batch;
        after_fused(loss): called once the loss is enqueued, before the weight-gradient kernel (bookkeeping on a side stream)."""
        obs, act, target = _rows_contig(_req(obs, "obs"), "obs"), _rows_contig(_req(act, "act"), "act"), _rows_contig(_req(target, "target"), "target")
        loss = torch.empty((self.E, self.B, self.head_dim), device=obs.device, dtype=torch.float32) if want_loss else None
        self.run_pack()
        if after_pack is not None:
            after_pack()
        self.run_fused(obs, act, target, mn, mx, reg, g_scale, loss)
        if after_fused is not None:
            after_fused(loss)
        self.run_dw()
        return loss


def fused_train_plan(model, E, B):
    """The model's FusedTrainPlan for batches of B rows per member, or None when the fused kernels do not take this model
    (deterministic head, CMI-test layout, per-sample masks, wide hidden layers ...: the per-layer kernels run instead)."""
    if not TRAIN_FUSED or model.deterministic or not hasattr(model, "hidden_layers") or getattr(model, "is_net_shard", False):
        return None  # (a net shard's outputs are not dimensions 0..P-1 of the observation: the per-layer kernels run)
    P = model._parallel_num()
    if not ((model._head_layout == HEAD_PLAIN and P == 1) or (model._head_layout == HEAD_PARALLEL and P == model._head_dim)):
        return None
    mask = model._kernel_mask()
    if mask is not None and not (mask.dim() == 2 or (mask.dim() == 3 and mask.shape[1] == E)):
        return None
    flat, flat_grad = model.flat_parameters()
    key = (E, flat.data_ptr(), flat_grad.data_ptr(), None if mask is None else (mask.data_ptr(), mask._version))
    plans = model._packed.get("train_plans")
    if plans is None or plans[0] != key:  # the arena or the mask moved: every plan points at stale memory
        plans = model._packed["train_plans"] = (key, {})
    plan = plans[1].get(B)
    if plan is None:
        if torch.cuda.is_current_stream_capturing():
            raise _lib.CmrlB200Error("the fused training plan must be built before the step is captured in a CUDA graph")
        if len(plans[1]) >= 8:  # ragged batch sizes of a growing data set: keep the set of live plans small
            for b_old, p_old in list(plans[1].items()):
                if not getattr(p_old, "pinned", False):  # a captured step graph holds the buffers of a pinned plan
                    del plans[1][b_old]
                    break
        plan = plans[1][B] = FusedTrainPlan(model, E, B)
    return plan if plan.ok else None
