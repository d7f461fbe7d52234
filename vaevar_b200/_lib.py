"""ctypes binding of libvaevar.so (include/vaevar.h).  No fallback: a missing library or a non-B200 device is an error."""
from __future__ import annotations

import ctypes as C
import pathlib

import os

_HERE = pathlib.Path(__file__).resolve().parent
# VV_LIB: an experimental build of the same library (tools/build_variant.sh) -- A/B measurements only
LIB_PATH = pathlib.Path(os.environ["VV_LIB"]).resolve() if os.environ.get("VV_LIB") else _HERE / "libvaevar.so"

VV_MAX_GROUPS = 8
VV_MAX_LG = 8


class NetConfigC(C.Structure):
    _fields_ = [("img_h", C.c_int), ("img_w", C.c_int), ("n_groups", C.c_int),
                ("in_chans", C.c_int * VV_MAX_GROUPS), ("out_chans", C.c_int * VV_MAX_GROUPS),
                ("enc_dim", C.c_int), ("embed_dim", C.c_int), ("window", C.c_int),
                ("enc_depth", C.c_int * 2), ("enc_heads", C.c_int * 2), ("n_lg", C.c_int),
                ("lg_depth", C.c_int * VV_MAX_LG), ("lg_heads", C.c_int * VV_MAX_LG), ("keep_out", C.c_int)]


class ConfigC(C.Structure):
    _fields_ = [("dec", NetConfigC), ("flow", NetConfigC), ("has_flow", C.c_int), ("T", C.c_int),
                ("recompute", C.c_int), ("use_graph", C.c_int), ("forward_fp16", C.c_int), ("no_ln_fold", C.c_int)]


VV_NET1_MAX_LEVELS = 4


class Net1ConfigC(C.Structure):
    _fields_ = [("img_h", C.c_int), ("img_w", C.c_int), ("n_groups", C.c_int),
                ("in_chans", C.c_int * VV_MAX_GROUPS), ("out_chans", C.c_int * VV_MAX_GROUPS),
                ("enc_dim", C.c_int), ("embed_dim", C.c_int), ("win_h", C.c_int), ("win_w", C.c_int), ("n_levels", C.c_int),
                ("enc_depth", C.c_int * VV_NET1_MAX_LEVELS), ("enc_heads", C.c_int * VV_NET1_MAX_LEVELS), ("n_lg", C.c_int),
                ("lg_depth", C.c_int * VV_MAX_LG), ("lg_heads", C.c_int * VV_MAX_LG), ("keep_out", C.c_int)]


_P = C.c_void_p
_SIGS = {
    "vv_last_error": (C.c_char_p, []),
    "vv_set_device": (C.c_int, [C.c_int]),
    "vv_engine_create": (C.c_int, [C.POINTER(ConfigC), C.POINTER(_P)]),
    "vv_engine_destroy": (None, [_P]),
    "vv_set_weight": (C.c_int, [_P, C.c_int, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "vv_finalize_weights": (C.c_int, [_P]),
    "vv_set_constants": (C.c_int, [_P, _P, _P, _P]),
    "vv_compact_mask": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, _P, C.POINTER(C.c_int64), _P]),
    "vv_set_case": (C.c_int, [_P, _P, _P, _P, _P, C.c_float, _P]),
    "vv_set_case_native": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_float, _P]),
    "vv_decode_native": (C.c_int, [_P, _P, _P, _P]),
    "vv_set_case_obsop": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_float, _P]),
    "vv_metrics": (C.c_int, [_P, _P, _P, _P, _P]),
    "vv_metrics_grid": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P]),
    "vv_num_obs": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "vv_ln_fold_health": (C.c_int, [_P, C.POINTER(C.c_uint32)]),
    "vv_cost_grad": (C.c_int, [_P, _P, _P, _P, _P]),
    "vv_cost": (C.c_int, [_P, _P, _P, _P]),
    "vv_decode": (C.c_int, [_P, _P, _P, _P]),
    "vv_integrate": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "vv_net_forward": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "vv_net_vjp": (C.c_int, [_P, C.c_int, _P, _P, _P, _P]),
    "vv_lbfgs_create": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P)]),
    "vv_lbfgs_destroy": (None, [_P]),
    "vv_lbfgs_reset": (C.c_int, [_P]),
    "vv_lbfgs_step": (C.c_int, [_P, _P, C.POINTER(C.c_double), _P]),
    "vv_lbfgs_history": (C.c_int, [_P, C.POINTER(C.c_double), C.c_int]),
    "vv_lbfgs_last_cost": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "vv_lbfgs_set_reuse": (C.c_int, [_P, C.c_int]),
    "vv_lbfgs_set_noise": (C.c_int, [_P, C.c_double]),
    "vv_lbfgs_steps": (C.c_int, [_P, C.POINTER(C.c_double), C.c_int]),
    "vv_lbfgs_create_testfn": (C.c_int, [C.c_longlong, C.c_int, C.c_int, C.POINTER(_P)]),
    "vv_profile_ops": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int]),
    "vv_test_gemm": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "vv_test_gemm_ln": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_float, _P, _P, _P, _P, C.POINTER(C.c_int), C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, _P, _P]),
    "vv_test_mlp_fwd": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P, _P, _P, _P]),
    "vv_test_mlp_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P]),
    "vv_test_lin_fwd": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P]),
    "vv_debug_mlp_trace": (C.c_int, [_P]),
    "vv_debug_cubic_interpolate": (C.c_double, [C.c_double] * 6 + [C.c_int] * 3 + [C.c_double] * 2),
    "vv_test_ln_stats": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "vv_debug_gemm_trace": (C.c_int, [_P]),
    "vv_debug_gemm_mode": (C.c_int, [C.c_int]),
    "vv_test_layernorm": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_float, _P]),
    "vv_test_winattn": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "vv_test_obs": (C.c_int, [_P, _P, _P, _P, _P]),
    "vv_last_launch_count": (C.c_int, [_P]),
    "vv_resample_nearest": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "vv_resample_nearest_adjoint": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "vv_obs_term_work_doubles": (C.c_int64, []),
    "vv_debug_seam_tables": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vv_obs_term": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, _P, _P, C.c_int64, _P, _P]),
    "vv_net1_create": (C.c_int, [C.POINTER(Net1ConfigC), C.POINTER(_P)]),
    "vv_net1_destroy": (None, [_P]),
    "vv_net1_set_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "vv_net1_finalize": (C.c_int, [_P]),
    "vv_net1_forward": (C.c_int, [_P, _P, _P, _P]),
    "vv_net1_set_constants": (C.c_int, [_P, _P, _P]),
    "vv_net1_integrate": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "vv_net1_last_launch_count": (C.c_int, [_P]),
    "vv_net1_device_bytes": (C.c_longlong, [_P]),
    "vv_net1_profile_ops": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, C.c_int]),
    "vv_test_attn1": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
}
EXPORTED = tuple(_SIGS)

_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library; raises if it has not been built (python -m vaevar_b200.build)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m vaevar_b200.build` "
                               "(there is no CPU / PyTorch fallback for this path)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class VVError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        raise VVError(load().vv_last_error().decode())
