"""ctypes binding of libdeeprest_b200.so (the C ABI in include/deeprest_b200.h).

There is no CPU fallback: if the shared object is missing or cannot be loaded
this module raises, and every compute entry point fails when no CUDA device is
present (``dr_create`` returns DR_ECUDA).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdeeprest_b200.so")

DR_OK, DR_EINVAL, DR_ECUDA, DR_ENOMEM, DR_ESTATE, DR_EUNSUPPORTED = 0, -1, -2, -3, -4, -5
ENGINE_AUTO, ENGINE_FFMA, ENGINE_TC = 0, 1, 2
DTYPES = {"fp32": 0, "f32": 0, "float32": 0, "bf16": 1, "bfloat16": 1}
ENGINES = {"auto": ENGINE_AUTO, "ffma": ENGINE_FFMA, "tcgen05": ENGINE_TC, "tc": ENGINE_TC}


class DrConfig(C.Structure):
    _fields_ = [
        ("F", C.c_int32), ("M", C.c_int32), ("H", C.c_int32), ("Q", C.c_int32),
        ("quantiles", C.c_float * 8), ("dropout_p", C.c_float),
        ("engine", C.c_int32), ("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
        ("dtype", C.c_int32),
    ]


# name -> (restype, argtypes); every symbol include/deeprest_b200.h declares
_FP = C.POINTER(C.c_float)
_H = C.c_void_p
SIGNATURES = {
    "dr_create": (C.c_int, [C.POINTER(DrConfig), C.POINTER(_H)]),
    "dr_destroy": (None, [_H]),
    "dr_last_error": (C.c_char_p, [_H]),
    "dr_version": (C.c_int, []),
    "dr_has_engine": (C.c_int, [C.c_int32]),
    "dr_set_stream": (C.c_int, [_H, C.c_void_p, C.c_int32]),
    "dr_profile": (C.c_int, [_H, C.c_int32]),
    "dr_profile_read": (C.c_int, [_H, C.POINTER(C.c_int32), _FP, _FP]),
    "dr_local_experts": (C.c_int, [_H, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "dr_load_weights": (C.c_int, [_H, _FP, C.c_size_t]),
    "dr_get_weights": (C.c_int, [_H, _FP, C.c_size_t]),
    "dr_forward": (C.c_int, [_H, _FP, C.c_int32, C.c_int32, _FP]),
    "dr_forward_dev": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "dr_series_windows": (C.c_int, [C.c_int32, C.c_int32, C.c_int32]),
    "dr_forward_series": (C.c_int, [_H, _FP, C.c_int32, C.c_int32, C.c_int32, _FP]),
    "dr_forward_series_dev": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "dr_set_output_transform": (C.c_int, [_H, _FP, _FP, C.c_float]),
    "dr_s_elems": (C.c_int64, [C.c_int32, C.c_int32]),
    "dr_forward_local_dev": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "dr_forward_heads_dev": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "dr_forward_heads_p2p_dev": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.c_int32, C.c_int64]),
    "dr_scatter_forecasts_dev": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.c_int32, C.c_int64]),
    "dr_interleave_dev": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "dr_comm_arena_bytes": (C.c_int64, [_H, C.c_int32, C.c_int32]),
    "dr_comm_init": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "dr_comm_attach": (C.c_int, [_H, C.c_void_p, C.POINTER(C.c_void_p)]),
    "dr_comm_detach": (C.c_int, [_H]),
    "dr_forward_sharded_dev": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "dr_forward_sharded_issue_dev": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]),
    "dr_forward_sharded_wait": (C.c_int, [_H, C.c_int32]),
    "dr_forward_sharded": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "dr_quantile_loss": (C.c_int, [_H, _FP, _FP, C.c_int32, C.c_int32, _FP]),
    "dr_quantile_loss_dev": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "dr_train_step": (C.c_int, [_H, _FP, _FP, C.c_int32, C.c_int32, C.c_void_p, C.c_uint64, C.c_float, _FP]),
    "dr_train_step_dev": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_uint64,
                                    C.c_float, C.c_void_p, C.c_void_p]),
    "dr_train_begin_dev": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_uint64,
                                     C.c_float, C.c_void_p, C.c_void_p]),
    "dr_train_advance": (C.c_int, [_H, C.POINTER(C.c_int32), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "dr_train_set_microbatch": (C.c_int, [_H, C.c_int32]),
    "dr_get_grads": (C.c_int, [_H, _FP, C.c_size_t]),
    "dr_debug_read": (C.c_int, [_H, C.c_char_p, _FP, C.c_size_t]),
    "dr_tc_probe": (C.c_int, [C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                              C.c_int32, C.c_int32, C.c_int32, _FP]),
    "dr_tc_probe_mn": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int32, C.c_int32,
                                 C.POINTER(C.c_uint32), _FP]),
    "dr_launch_count": (C.c_int64, [_H]),
    "dr_last_engine": (C.c_char_p, [_H]),
}

_lib = None


class DeepRestError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libdeeprest_b200 error {code}: {msg}")
        self.code = code


def load():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing — build it with `python -m deeprest_b200.build` "
                "(there is no CPU fallback for the estimator hot path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(handle, rc):
    if rc != DR_OK:
        msg = load().dr_last_error(handle)
        raise DeepRestError(rc, msg.decode() if msg else "?")
