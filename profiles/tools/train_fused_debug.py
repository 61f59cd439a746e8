"""Stage-by-stage check of the fused tcgen05 training step (train_tc.cu) against an fp64 torch reference computed on the
CPU: every saved tile (layer inputs, pre-activations, dz), the loss and every gradient; then kernel timings.
Usage: python profiles/tools/train_fused_debug.py [--quick]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cmrl_b200 as cm  # noqa: E402
from cmrl_b200 import ops  # noqa: E402

SILU = {"_target_": "torch.nn.SiLU"}
dev = "cuda:0"


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def tile_to_rows(flat, cols):  # [row quarter][col / 4][32 rows][4]
    return flat.reshape(4, cols // 4, 32, 4).transpose(0, 2, 1, 3).reshape(128, cols)


def reference(model, obs, act, target, g_scale):
    """fp64 forward / backward of every net; returns per-net dicts of intermediates and parameter gradients."""
    E = model.ensemble_num
    P = model._parallel_num()
    params = [p.detach().double().cpu().requires_grad_(True) for p in model._stack_params()]
    L = model.num_layers
    mn, mx = params[-2], params[-1]
    x = torch.cat([obs, act], -1).double().cpu()  # [E,B,in]
    mask = model._kernel_mask()
    parallel = params[0].dim() == 4
    if parallel:
        xin = x.unsqueeze(0).expand(P, -1, -1, -1)
        if mask is not None:
            m = mask.double().cpu()
            xin = xin * (m[:, None, None, :] if m.dim() == 2 else m[:, :, None, :])
    else:
        xin = x
    h, zs, hs = xin, [], [xin]
    for l in range(L):
        z = h @ params[2 * l] + params[2 * l + 1]
        z.retain_grad()
        zs.append(z)
        h = torch.nn.functional.silu(z) if model._act_kind == ops.ACT_SILU else (torch.relu(z) if model._act_kind == ops.ACT_RELU else z)
        hs.append(h)
    raw = h @ params[2 * L] + params[2 * L + 1]
    raw.retain_grad()
    sp = torch.nn.functional.softplus
    if parallel:
        mean = raw[..., 0].permute(1, 2, 0)
        lvr = raw[..., 1]
        lv = mx.view(P, 1, 1) - sp(mx.view(P, 1, 1) - lvr)
        lv = (mn.view(P, 1, 1) + sp(lv - mn.view(P, 1, 1))).permute(1, 2, 0)
    else:
        D = model._head_dim
        mean, lvr = raw[..., :D], raw[..., D:]
        lv = mx - sp(mx - lvr)
        lv = mn + sp(lv - mn)
    if model._residual():
        mean = mean + obs.double().cpu()
    t = target.double().cpu()
    loss = (mean - t) ** 2 * torch.exp(-lv) + lv + 0.01 * (mx.sum() - mn.sum()).detach()
    (loss.sum() * g_scale).backward()
    return dict(loss=loss.detach().numpy(), zs=[z.detach().numpy() for z in zs], hs=[h.detach().numpy() for h in hs],
                dzs=[z.grad.numpy() for z in zs] + [raw.grad.numpy()], grads=[p.grad for p in params[:2 * (L + 1)]], x=x.numpy())


def check(name, model, B, seed=0):
    torch.manual_seed(seed)
    E, O, A = model.ensemble_num, model.obs_size, model.action_size
    with torch.no_grad():  # non-trivial biases
        for p in model.parameters():
            if p.dim() >= 3 and p.shape[-2] == 1:
                p.add_(0.1 * torch.randn_like(p))
    obs = torch.randn(E, B, O, device=dev)
    act = torch.rand(E, B, A, device=dev) * 2 - 1
    D = model._head_dim
    target = obs[..., :D] + 0.1 * torch.randn(E, B, D, device=dev)
    g_scale = 1.0 / (E * B * D)
    model.flat_parameters()
    plan = ops.fused_train_plan(model, E, B)
    assert plan is not None, "model not taken by the fused path"
    mn, mx = model._bounds()
    loss = plan.step(obs, act, target, mn, mx, model._nll_reg(), g_scale)
    torch.cuda.synchronize()
    ref = reference(model, obs.cpu(), act.cpu(), target.cpu(), g_scale)
    print("== %s  B=%d  E=%d P=%d H=%d L=%d" % (name, B, E, plan.P, plan.hidden, plan.L))
    print("   loss rel err %.2e" % rel(loss.cpu().numpy(), ref["loss"]))
    # saved tiles of net 0 / last net, tile 0
    import ctypes
    blob, scratch, ws = ctypes.c_longlong(0), ctypes.c_longlong(0), ctypes.c_longlong(0)
    cm._lib.load().cmrl_train_fused_sizes(plan.L, plan.in_size, plan.hidden, plan.head_outputs, plan.P, E, B,
                                          ctypes.byref(blob), ctypes.byref(scratch), ctypes.byref(ws))
    tiles = (B + 127) // 128
    slots = (tiles + 1) // 2 * 2  # tile slots per net are padded to cluster pairs
    item = scratch.value // (plan.P * E * slots)
    sc = plan.scratch.cpu().numpy()
    K0P = (plan.in_size + 1 + 15) // 16 * 16
    HP = (plan.hidden + 1 + 15) // 16 * 16
    NH = (2 * plan.head_outputs + 15) // 16 * 16
    H, L = plan.hidden, plan.L
    worst = {}
    for net in sorted({0, plan.P * E - 1, (plan.P * E) // 2}):
        p_, e_ = net // E, net % E
        for tile in sorted({0, tiles - 1}):
            base = (net * slots + tile) * item
            r0, nr = tile * 128, min(128, B - tile * 128)
            off = 0
            x = tile_to_rows(sc[base + off:base + off + K0P * 128], K0P); off += K0P * 128
            worst["x"] = max(worst.get("x", 0), rel(x[:nr, :plan.in_size], ref["x"][e_, r0:r0 + nr]))
            sel = (lambda a: a[p_, e_]) if plan.P > 1 or ref["zs"][0].ndim == 4 else (lambda a: a[e_])
            for l in range(L):
                z = tile_to_rows(sc[base + off:base + off + HP * 128], HP); off += HP * 128
                worst["z%d" % l] = max(worst.get("z%d" % l, 0), rel(z[:nr, :H], sel(ref["zs"][l])[r0:r0 + nr]))
            for l in range(L):
                dz = tile_to_rows(sc[base + off:base + off + HP * 128], HP); off += HP * 128
                worst["dz%d" % l] = max(worst.get("dz%d" % l, 0), rel(dz[:nr, :H], sel(ref["dzs"][l])[r0:r0 + nr]))
            dzh = tile_to_rows(sc[base + off:base + off + NH * 128], NH)
            Dn = plan.head_outputs
            dzh_ref = sel(ref["dzs"][L])[r0:r0 + nr]
            got = np.stack([dzh[:nr, 0:2 * Dn:2], dzh[:nr, 1:2 * Dn:2]], 1).reshape(nr, 2 * Dn) if Dn > 1 else dzh[:nr, :2]
            worst["dz_head"] = max(worst.get("dz_head", 0), rel(got, dzh_ref))
    print("   saved tiles (max rel err):", "  ".join("%s %.1e" % kv for kv in worst.items()))
    params = model._stack_params()
    gerr = [rel(params[i].grad.cpu().numpy(), ref["grads"][i].numpy()) for i in range(2 * (L + 1))]
    print("   grads  dW/db per layer:", "  ".join("%.1e/%.1e" % (gerr[2 * l], gerr[2 * l + 1]) for l in range(L + 1)))
    return max([rel(loss.cpu().numpy(), ref["loss"])] + gerr)


def timing(B, kind="plain"):
    if kind == "plain":
        model = cm.PlainTransition(5, 1, hid_size=200, activation_fn_cfg=SILU, device=dev)
    else:
        model = cm.ExternalMaskTransition(5, 1, hid_size=200, activation_fn_cfg=SILU, device=dev)
    E = 7
    obs = torch.randn(E, B, 5, device=dev)
    act = torch.rand(E, B, 1, device=dev)
    target = obs + 0.1 * torch.randn(E, B, 5, device=dev)
    model.flat_parameters()
    plan = ops.fused_train_plan(model, E, B)
    mn, mx = model._bounds()
    import ctypes
    lib = cm._lib
    for passes in (3, 1):
        ops.TRAIN_FUSED_PASSES = passes
        for _ in range(3):
            plan.step(obs, act, target, mn, mx, 0.5, 1e-4)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        reps = 20
        ev[0].record()
        for _ in range(reps):
            plan.step(obs, act, target, mn, mx, 0.5, 1e-4)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        flop = 736800.0 * E * B * (5 if kind != "plain" else 1)
        print("timing %s B=%d passes=%d: pack+fused+dw %.4f ms  (%.1f TFLOP/s algorithmic)" % (kind, B, passes, ms, flop / ms / 1e9))
    ops.TRAIN_FUSED_PASSES = 3
    # the three launches separately
    st = ops._stream()
    def t(fn, reps=20):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    m = plan.mask
    pk = lambda: lib.call("cmrl_train_pack_tf32", plan.w_ptrs, plan.b_ptrs, plan.L, plan.P, E, plan.in_size, plan.hidden, plan.head_outputs,
                          ops._ptr(m), plan.mask_str[0], plan.mask_str[1], plan.packed.data_ptr(), st)
    fu = lambda: lib.call("cmrl_train_fused", plan.packed.data_ptr(), plan.L, plan.P, E, B, 5, 1, plan.hidden, plan.head_dim, plan.layout,
                          plan.act, plan.residual, 3, obs.data_ptr(), obs.stride(0), act.data_ptr(), act.stride(0), target.data_ptr(),
                          target.stride(0), 5, mn.data_ptr(), mx.data_ptr(), mn.numel(), 0.5, 1e-4, None, plan.scratch.data_ptr(), st)
    dw = lambda: lib.call("cmrl_train_dw", plan.scratch.data_ptr(), plan.L, plan.P, E, B, plan.in_size, plan.hidden, plan.head_outputs,
                          plan.act, plan.dw_ptrs, plan.db_ptrs, ops._ptr(m), plan.mask_str[0], plan.mask_str[1], ops._ptr(plan.dw_ws),
                          0 if plan.dw_ws is None else plan.dw_ws.numel(), st)
    print("   alone: pack %.4f ms, fused %.4f ms, dw %.4f ms" % (t(pk), t(fu), t(dw)))


if __name__ == "__main__":
    worst = 0.0
    worst = max(worst, check("plain tiny", cm.PlainTransition(5, 1, hid_size=24, activation_fn_cfg=SILU, device=dev), 32))
    worst = max(worst, check("plain C1 256", cm.PlainTransition(5, 1, hid_size=200, activation_fn_cfg=SILU, device=dev), 256))
    worst = max(worst, check("plain C1 ragged 300", cm.PlainTransition(5, 1, hid_size=200, activation_fn_cfg=SILU, device=dev), 300))
    worst = max(worst, check("reward relu", cm.PlainRewardMech(5, 1, hid_size=64, device=dev), 130))
    m = cm.ExternalMaskTransition(5, 1, hid_size=40, activation_fn_cfg=SILU, device=dev)
    mask = (torch.rand(5, 6) < 0.5).float()
    mask[torch.arange(5), torch.arange(5)] = 1
    m.set_input_mask(mask)
    worst = max(worst, check("extmask 2d", m, 100))
    worst = max(worst, check("plain 2048 rows (split dw)", cm.PlainTransition(5, 1, hid_size=200, activation_fn_cfg=SILU, device=dev), 2048))
    print("WORST rel err %.2e" % worst)
    if "--quick" not in sys.argv:
        for B in (256, 4096, 16384):
            timing(B)
        timing(256, "extmask")
        timing(2048, "extmask")
