"""ctypes binding of include/qeft_b200.h.  Fails loudly when the CUDA library is unavailable."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# (QEFT_B200_LIB: development switch, an alternative build of the same library for A/B timing -- tools/build_variant.sh)
LIB_PATH = os.environ.get("QEFT_B200_LIB") or os.path.join(_HERE, "csrc", "libqeft_b200.so")

OK = 0
E_UNSUPPORTED = -6
DT_F16, DT_BF16 = 0, 1
OW_NONE, OW_PLAIN, OW_INTERLEAVED = 0, 1, 2
F_PDL = 1
GEMV_MAX_PARTS = 4


class GemvPart(C.Structure):
    _fields_ = [("qweight", C.c_void_p), ("scales", C.c_void_p), ("scaled_zeros", C.c_void_p),
                ("oweight", C.c_void_p), ("bias", C.c_void_p), ("y", C.c_void_p), ("N", C.c_int)]


MAX_RANKS = 8


class Gather(C.Structure):
    _fields_ = [("nranks", C.c_int), ("y_ld", C.c_int),
                ("y_peer", (C.c_void_p * GEMV_MAX_PARTS) * MAX_RANKS),
                ("done_peer", C.c_void_p * MAX_RANKS),
                ("wait_flag", C.c_void_p), ("epoch", C.c_void_p),
                ("y_mc", C.c_void_p * GEMV_MAX_PARTS)]


EPI_NONE, EPI_SWIGLU, EPI_RESIDUAL = 0, 1, 2


class DecodeStage(C.Structure):
    """qeft_decode_stage_t (include/qeft_b200.h)."""
    _fields_ = [("parts", GemvPart * GEMV_MAX_PARTS), ("nparts", C.c_int), ("K", C.c_int), ("r", C.c_int),
                ("G", C.c_int), ("x", C.c_void_p), ("x_gather", C.c_void_p), ("norm_weight", C.c_void_p),
                ("norm_eps", C.c_float), ("epilogue", C.c_int), ("residual", C.c_void_p)]


_vp, _i, _u = C.c_void_p, C.c_int, C.c_uint
# name -> (restype, argtypes): every symbol include/qeft_b200.h declares
SIGNATURES = {
    "qeft_abi_version": (_i, []),
    "qeft_build_info": (C.c_char_p, []),
    "qeft_launch_count": (C.c_uint64, []),
    "qeft_status_string": (C.c_char_p, [_i]),
    "qeft_gemv_w4": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _u, _vp]),
    "qeft_gemv_w4_multi": (_i, [_vp, C.POINTER(GemvPart), _i, _i, _vp, _i, _i, _i, _i, _u, _vp]),
    "qeft_gemv_w4_multi_gather": (_i, [_vp, C.POINTER(GemvPart), _i, _i, _vp, _i, _i, _i, _i, _u, C.POINTER(Gather), _vp]),
    "qeft_gather_wait": (_i, [_vp, _vp, _i, _vp]),
    "qeft_decode_program_create": (_i, [C.POINTER(DecodeStage), _i, _i, C.POINTER(C.c_void_p)]),
    "qeft_decode_program_run": (_i, [_vp, _i, _i, _u, _vp]),
    "qeft_decode_program_num_stages": (_i, [_vp]),
    "qeft_decode_program_set_ranks": (_i, [_vp, _i, _i, C.POINTER(C.c_void_p)]),
    "qeft_decode_program_shard": (_i, [_vp, _i, _i, C.POINTER(C.c_void_p)]),
    "qeft_decode_program_destroy": (_i, [_vp]),
    "qeft_gemm_w4": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _u, _vp]),
    "qeft_gemm_w4_gather": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _u, C.POINTER(Gather), _vp]),
    "qeft_gemm_w4_dx": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _u, _vp]),
    "qeft_gemm_w4_dx_plan": (_i, [_i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "qeft_dow": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _u, _vp]),
    "qeft_pack_w4": (_i, [_vp, _vp, _i, _i, _vp]),
    "qeft_unpack_w4": (_i, [_vp, _vp, _i, _i, _vp]),
    "qeft_dequant_w4": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "qeft_interleave_oweight": (_i, [_vp, _vp, _i, _i, _i, _vp]),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load (building first if the in-tree .so is missing or stale and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH) or os.environ.get("QEFT_B200_REBUILD") == "1":
            from . import build as _build
            _build.build(force=os.environ.get("QEFT_B200_REBUILD") == "1")
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"qeft_b200: CUDA library {LIB_PATH} is missing and could not be built; "
                               "there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError if the header and the library diverge
            fn.restype = res
            fn.argtypes = args
        if lib.qeft_abi_version() != 1:
            raise RuntimeError("qeft_b200: ABI version mismatch between _lib.py and libqeft_b200.so")
        _lib = lib
    return _lib


def check(status: int, what: str):
    if status != OK:
        msg = load().qeft_status_string(status).decode()
        raise RuntimeError(f"{what}: {msg} (status {status})")


def launch_count() -> int:
    return int(load().qeft_launch_count())
