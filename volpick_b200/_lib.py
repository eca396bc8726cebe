"""ctypes binding of ``libvolpick_b200.so`` (the C ABI declared in include/volpick_b200.h).

The library is built in-tree by ``python -m volpick_b200.build`` / ``__graft_entry__.build()``.
There is no fallback: if the shared object is missing, ``load()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VP_LIB_PATH") or os.path.join(HERE, "libvolpick_b200.so")  # VP_LIB_PATH: A/B builds

VP_OK = 0
VP_ERR_ARG, VP_ERR_CUDA, VP_ERR_CAPACITY, VP_ERR_WORKSPACE, VP_ERR_UNSUPPORTED = -1, -2, -3, -4, -5
KIND_EQTRANSFORMER, KIND_PHASENET = 0, 1
STACK = {"avg": 0, "max": 1}
PRECISION = {"fp32": 0, "f16x3": 1, "bf16": 2}
DTYPE_F32, DTYPE_I32 = 0, 1
# vp_trigger as a NumPy record (include/volpick_b200.h)
import numpy as _np  # noqa: E402

TRIGGER_DTYPE = _np.dtype([("s0", "<i8"), ("s1", "<i8"), ("s_peak", "<i8"), ("value", "<f4"), ("label", "<i4")])
PEAK_SCOPE = {"channel": 0, "window": 1}
PRE_TAPER, PRE_DETREND = 1, 2


class Trigger(C.Structure):
    """``vp_trigger``"""

    _fields_ = [("s0", C.c_int64), ("s1", C.c_int64), ("s_peak", C.c_int64), ("value", C.c_float), ("label", C.c_int32)]


class AnnotateParams(C.Structure):
    """``vp_annotate_params``"""

    _fields_ = [
        ("overlap", C.c_int64),
        ("blinding", C.c_int64 * 2),
        ("stacking", C.c_int32),
        ("precision", C.c_int32),
        ("peak_scope", C.c_int32),
        ("chunk_windows", C.c_int32),
        ("threshold", C.c_float * 3),
        ("norm_detrend", C.c_int32),
    ]


# name -> (restype, argtypes); every symbol declared in include/volpick_b200.h
_vp = C.c_void_p
_i64, _i32, _f32 = C.c_int64, C.c_int, C.c_float
SIGNATURES = {
    "vp_version": (_i32, []),
    "vp_last_error": (C.c_char_p, []),
    "vp_launch_count": (_i64, [_i32]),
    "vp_kernel_timing": (_i32, [_i32]),
    "vp_kernel_class_names": (C.c_char_p, []),
    "vp_kernel_timing_read": (_i32, [_i32, C.POINTER(C.c_double), C.POINTER(_i64)]),
    "vp_model_create": (_i32, [_i32, _vp, _i64, _i32, C.POINTER(_vp)]),
    "vp_model_destroy": (_i32, [_vp]),
    "vp_model_kind": (_i32, [_vp]),
    "vp_model_in_samples": (_i32, [_vp]),
    "vp_model_expected_floats": (_i64, [_i32]),
    "vp_window_count": (_i64, [_i64, _i64, _i64]),
    "vp_window_starts": (_i32, [_i64, _i64, _i64, _vp, _i64, C.POINTER(_i64)]),
    "vp_coverage": (_i64, [_i64, _i64]),
    "vp_sosfilt_workspace_bytes": (_i64, [_i64, _i32, _i32]),
    "vp_sosfilt": (_i32, [_vp, _i32, _i64, _i64, _i32, _vp, _i32, _i32, _vp, _vp, _i64, _vp]),
    "vp_slice_normalize": (_i32, [_vp, _i32, _i64, _i64, _vp, _i64, _i64, _i32, _i32, _vp, _vp]),
    "vp_forward_workspace_bytes": (_i64, [_vp, _i64, _i32]),
    "vp_forward": (_i32, [_vp, _vp, _i64, _vp, _vp, _i64, _i32, _vp]),
    "vp_forward_range": (_i32, [_vp, _vp, _i64, _vp, _vp, _i64, _i32, _i64, _i64, _vp]),
    "vp_slice_forward": (_i32, [_vp, _vp, _i32, _i64, _i64, _vp, _i64, _i32, _i32, _vp, _vp, _i64, _i32, _i64, _i64, _vp]),
    "vp_forward_tap": (_i32, [_vp, _vp, _i64, _vp, _vp, _i64, _i32, C.c_char_p, _vp, _i64, C.POINTER(_i64), _vp]),
    "vp_forward_tap_names": (C.c_char_p, [_vp]),
    "vp_tcconv_debug": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "vp_tcconv_bench": (_i32, [_i32] * 11 + [C.POINTER(C.c_float)]),
    "vp_stack": (_i32, [_vp, _vp, _i64, _i64, _i32, _i64, _i64, _i64, _i32, _vp, _i64, _vp]),
    "vp_nan_bounds": (_i32, [_vp, _i32, _i64, _vp, _vp]),
    "vp_pick_scratch_bytes": (_i64, [_i64]),
    "vp_pick": (_i32, [_vp, _i64, _f32, _f32, _i32, _vp, _i64, _vp, _vp, _i64, _vp]),
    "vp_pick_labels": (_i32, [_vp, _i32, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "vp_pick_windows": (_i32, [_vp, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "vp_annotate_workspace_bytes": (_i64, [_vp, _i64, C.POINTER(AnnotateParams), _i32, _i64]),
    "vp_annotate": (_i32, [_vp, _vp, _i32, _i32, _i64, _i64, C.POINTER(AnnotateParams), _vp, _i32, _vp, _i64,
                           C.POINTER(_i64), _vp, _vp, _i64, _vp]),
    "vp_annotate_begin": (_i32, [_vp, _vp, _i32, _i32, _i64, _i64, C.POINTER(AnnotateParams), _vp, _i32, _i64, _vp, _i64, _vp,
                                 C.POINTER(_vp)]),
    "vp_annotate_end": (_i32, [_vp, _vp, _i64, C.POINTER(_i64), _vp]),
}

_lib: Optional[C.CDLL] = None


class VolpickError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"volpick_b200 error {code}: {message}")
        self.code = code


def load() -> C.CDLL:
    """Load the CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m volpick_b200.build` "
                "(there is no CPU or PyTorch fallback for the picking path)"
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(code: int) -> int:
    if code < 0:
        raise VolpickError(code, load().vp_last_error().decode(errors="replace"))
    return code
