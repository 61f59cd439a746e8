"""ctypes binding of libcmrl_b200.so (the C ABI declared in include/cmrl_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails the
error propagates.  The library is built in-tree by `csrc/build.py` (nvcc, sm_100a).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcmrl_b200.so")

ACT_IDENTITY, ACT_RELU, ACT_SILU = 0, 1, 2
HEAD_PLAIN, HEAD_PARALLEL = 0, 1

_c = ctypes
_P = _c.c_void_p  # every device pointer / stream crosses the ABI as void*
_I = _c.c_int
_LL = _c.c_longlong
_ULL = _c.c_ulonglong
_F = _c.c_float

# name -> argtypes (mirrors include/cmrl_b200.h one to one; tests check the symbols exist)
PROTOTYPES = {
    "cmrl_ens_linear_fwd": [_P, _LL, _LL, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _LL, _LL, _LL, _P, _I, _I, _P],
    "cmrl_ens_linear_bwd_dx": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "cmrl_ens_linear_bwd_dw": [_P, _LL, _LL, _P, _LL, _LL, _LL, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _LL, _P],
    "cmrl_gaussian_head_fwd": [_P, _I, _I, _I, _I, _P, _LL, _I, _P, _P, _I, _P, _P, _P],
    "cmrl_gaussian_head_bwd": [_P, _I, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P],
    "cmrl_nll_loss_fwd": [_P, _P, _P, _LL, _I, _LL, _F, _P, _P],
    "cmrl_nll_loss_bwd": [_P, _P, _P, _LL, _I, _LL, _P, _F, _P, _P, _P],
    "cmrl_mse_loss_fwd": [_P, _P, _LL, _I, _LL, _P, _P],
    "cmrl_mse_loss_bwd": [_P, _P, _LL, _I, _LL, _P, _P, _P],
    "cmrl_mse_member_sums": [_P, _P, _LL, _I, _LL, _P, _P],
    "cmrl_adam_l2_step": [_P, _P, _P, _P, _LL, _F, _F, _F, _F, _F, _F, _I, _P],
    "cmrl_adam_l2_step_dev": [_P, _P, _P, _P, _LL, _F, _F, _F, _F, _F, _F, _P, _P],
    "cmrl_sum_acc_f64": [_P, _LL, _P, _P, _P],
    "cmrl_bootstrap_gather": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "cmrl_draw_member_noise": [_ULL, _ULL, _P, _I, _P, _I, _I, _P, _P, _P],
    "cmrl_draw_uniform_index": [_ULL, _ULL, _P, _I, _I, _P, _P],
    "cmrl_counter_add": [_P, _ULL, _P],
    "cmrl_member_bucket": [_P, _I, _I, _P, _P, _P, _P],
    "cmrl_gather_inputs": [_P, _P, _P, _I, _I, _I, _P, _P],
    "cmrl_sample_scatter": [_P, _I, _P, _I, _I, _P, _I, _P, _P, _I, _P, _P, _P],
    "cmrl_select_sample": [_P, _P, _P, _I, _I, _I, _P, _P, _P],
    "cmrl_mopo_penalty": [_P, _I, _I, _I, _P, _P],
    "cmrl_pack_layer_f16": [_P, _P, _P, _LL, _LL, _I, _I, _I, _I, _I, _I, _I, ctypes.c_float, _P, _LL, _P],
    "cmrl_mlp_rollout_f16": [_P, _LL, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P, _P,
                             _I, _P, _P, _LL, _I, _P],
    "cmrl_mlp_rollout_f16_groups": [_P, _I, _P],
    "cmrl_train_pack_tf32": [_P, _P, _I, _I, _I, _I, _I, _I, _P, _LL, _LL, _P, _P],
    "cmrl_train_fused": [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _LL, _P, _LL, _P, _LL, _I, _P, _P, _I, _F, _F,
                         _P, _P, _P],
    "cmrl_mlp_rollout_tf32": [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _P, _P, _LL, _P],
    "cmrl_train_dw": [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _LL, _LL, _P, _LL, _P],
    "cmrl_env_step_tail": [_P, _P, _P, _P, _P, _F, _P, _I, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    "cmrl_dp_comm_create": [_I, _I, _LL, _c.POINTER(_P), _c.c_char_p],
    "cmrl_dp_comm_connect": [_P, _c.c_char_p],
    "cmrl_dp_comm_error": [_P, _c.POINTER(_I)],
    "cmrl_dp_comm_destroy": [_P],
    "cmrl_allreduce_grads": [_P, _P, _LL, _P],
    "cmrl_allreduce_grads_adam": [_P, _P, _P, _P, _P, _LL, _F, _F, _F, _F, _F, _F, _P, _P],
}

class RolloutGroup(_c.Structure):
    """struct cmrl_rollout_group (include/cmrl_b200.h): one mechanism of a grouped rollout launch."""

    _fields_ = [("packed", _P), ("blob_halfs", _LL), ("layer_off_halfs", _P), ("layer_k", _P), ("layer_n", _P)] + \
               [(n, _I) for n in ("num_layers", "P", "E", "N", "O", "A", "H", "head_dim", "layout", "gaussian", "act", "residual", "bounds_n")] + \
               [(n, _P) for n in ("obs", "actn", "perm", "offsets", "min_logvar", "max_logvar", "noise", "pred")]


_lib = None


class CmrlB200Error(RuntimeError):
    pass


def load():
    """Load the shared library once; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CmrlB200Error(
            "libcmrl_b200.so not found at %s -- build it with `python causal-mbrl_b200/csrc/build.py` "
            "(there is no CPU / PyTorch fallback)" % LIB_PATH
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.cmrl_abi_version.restype = _I
    lib.cmrl_last_error.restype = _c.c_char_p
    lib.cmrl_launch_count.restype = _LL
    lib.cmrl_train_fused_sizes.restype = _LL
    lib.cmrl_train_fused_sizes.argtypes = [_I, _I, _I, _I, _I, _I, _I, _c.POINTER(_LL), _c.POINTER(_LL), _c.POINTER(_LL)]
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _I
    _lib = lib
    return lib


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.cmrl_last_error().decode("utf-8", "replace")
        raise CmrlB200Error("%s failed (code %d): %s" % (name, rc, msg))


_replayed = 0


def note_replayed(n):
    """Kernels launched by replaying a captured CUDA graph (the C counter only sees them once, at capture)."""
    global _replayed
    _replayed += int(n)


def launch_count():
    """Kernels of this library launched so far: direct launches (counted in C) + graph-replayed ones."""
    return int(load().cmrl_launch_count()) + _replayed
