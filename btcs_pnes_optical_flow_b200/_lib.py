"""ctypes binding of libbtcsflow.so (include/btcsflow.h).  No CPU fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from .build import LIB_PATH

OPTFLOW_USE_INITIAL_FLOW = 4
OPTFLOW_FARNEBACK_GAUSSIAN = 256
BF_E_INVALID, BF_E_UNSUPPORTED, BF_E_NODEVICE = -1, -2, -3
BF_DTYPE_U8, BF_DTYPE_F32 = 0, 1
BF_PLAN_EXACT_F32 = 1
BF_PROF_TAGS = ("iter_update", "iter_last", "update", "coarse", "expand")   # BF_PROF_* in include/btcsflow.h


class BfParams(C.Structure):
    _fields_ = [
        ("pyr_scale", C.c_double),
        ("levels", C.c_int),
        ("winsize", C.c_int),
        ("iterations", C.c_int),
        ("poly_n", C.c_int),
        ("poly_sigma", C.c_double),
        ("flags", C.c_int),
    ]


class BtcsFlowError(RuntimeError):
    """Raised for every non-zero return of the C library (code, message from bf_last_error)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"btcsflow error {code}: {msg}")
        self.code = code
        self.msg = msg


class Cv2CompatError(ValueError, BtcsFlowError):
    """Invalid-argument errors that cv2 reports as cv2.error (-215 assertion)."""

    def __init__(self, code: int, msg: str):
        BtcsFlowError.__init__(self, code, msg)


_vp, _i, _d, _sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
_SIGNATURES = {
    "bf_plan_create": (_i, [C.POINTER(BfParams), _i, _i, _i, _i, _i, C.POINTER(_vp)]),
    "bf_plan_create_ex": (_i, [C.POINTER(BfParams), _i, _i, _i, _i, _i, C.c_uint, C.POINTER(_vp)]),
    "bf_plan_destroy": (_i, [_vp]),
    "bf_plan_coeff_storage": (_i, [_vp]),
    "bf_plan_workspace_bytes": (_sz, [_vp]),
    "bf_plan_num_scales": (_i, [_vp]),
    "bf_plan_scale_info": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_d), C.POINTER(_i)]),
    "bf_plan_profile": (_i, [_vp, _i]),
    "bf_plan_profile_read": (_i, [_vp, C.POINTER(_i), C.POINTER(_d), C.POINTER(C.c_longlong)]),
    "bf_plan_profile_tag": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_d), C.POINTER(C.c_longlong)]),
    "bf_launch_count": (C.c_longlong, []),
    "bf_launch_count_reset": (None, []),
    "bf_flow_pair": (_i, [_vp, _vp, _vp, _i, _sz, _vp, _vp]),
    "bf_flow_pair_host": (_i, [_vp, _vp, _vp, _i, _sz, _vp, _vp]),
    "bf_flow_series": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "bf_flow_series_host": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "bf_flow_series_host_async": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, C.POINTER(C.c_longlong)]),
    "bf_flow_series_wait": (_i, [_vp, C.c_longlong]),
    "bf_pc1_sliding": (_i, [_vp, _vp, _i, _i, _i, _d, _d, _i, _vp, _vp]),
    "bf_pc1_sliding_batched": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _d, _d, _i, _vp, _vp]),
    "bf_pc1_sliding_host": (_i, [_vp, _vp, _i, _i, _i, _d, _d, _i, _vp]),
    "bf_bgr2gray": (_i, [_vp, _i, _i, _i, _sz, _vp, _sz, _vp]),
    "bf_bandpass_nanrobust": (_i, [_vp, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "bf_sosfilt_zi": (_i, [_vp, _i, _vp]),
    "bf_stage_level_image": (_i, [_vp, _vp, _i, _sz, _i, _vp, _vp]),
    "bf_stage_poly_exp": (_i, [_vp, _i, _i, _i, _d, _vp, _vp]),
    "bf_stage_update_matrices": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "bf_stage_blur_solve": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "bf_stage_upsample_flow": (_i, [_vp, _i, _i, _i, _i, C.c_float, _vp, _vp]),
    "bf_last_error": (C.c_char_p, []),
    "bf_device_sm": (_i, [_i]),
    "bf_version": (C.c_char_p, []),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def lib_path() -> Path:
    return LIB_PATH


def load() -> C.CDLL:
    """Load the in-tree shared library; never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m btcs_pnes_optical_flow_b200.build` "
            "(or __graft_entry__.build()).  There is no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc == 0:
        return
    msg = load().bf_last_error().decode("utf-8", "replace")
    if rc == BF_E_INVALID:
        raise Cv2CompatError(rc, msg)
    raise BtcsFlowError(rc, msg)
