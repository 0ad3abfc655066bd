"""ctypes binding of libdeo_b200.so (include/deo_b200.h).  The library is the product; there is no
CPU fallback: if the shared object is missing, or no CUDA device is present, calls fail loudly."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DEO_LIB_PATH") or os.path.join(_HERE, "libdeo_b200.so")   # override: A/B builds

DEO_OK, DEO_ERR_INVALID, DEO_ERR_CUDA, DEO_ERR_UNSUPPORTED, DEO_ERR_NCCL, DEO_ERR_NOMEM = range(6)
DEO_F32, DEO_F64 = 0, 1
DEO_OP_CENTERED, DEO_OP_UPWIND = 0, 1
DEO_BC_NONE, DEO_BC_AFFINE, DEO_BC_PERIODIC = 0, 1, 2
DEO_FLAG_FORCE_GENERIC = 1
DEO_MAX_DIMS, DEO_MAX_OPS, DEO_MAX_TAPS = 3, 16, 17
DEO_DIST_ID_BYTES = 128


class DeoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libdeo_b200 error {code}: {msg}")
        self.code = code


class OpDesc(C.Structure):
    _fields_ = [("axis", C.c_int32), ("kind", C.c_int32), ("nonuniform", C.c_int32),
                ("derivative_order", C.c_int32), ("len", C.c_int32), ("stencil_length", C.c_int32),
                ("boundary_stencil_length", C.c_int32), ("boundary_point_count", C.c_int32),
                ("offside", C.c_int32), ("reserved", C.c_int32),
                ("stencil_coefs", C.c_void_p), ("low_boundary_coefs", C.c_void_p),
                ("high_boundary_coefs", C.c_void_p), ("coefficients", C.c_void_p)]


class BcDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("per_face", C.c_int32), ("K_l", C.c_int32), ("K_r", C.c_int32),
                ("a_l", C.c_void_p), ("b_l", C.c_void_p), ("a_r", C.c_void_p), ("b_r", C.c_void_p)]


class PlanDesc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("ndims", C.c_int32), ("dims", C.c_int64 * DEO_MAX_DIMS),
                ("padded", C.c_int32 * DEO_MAX_DIMS), ("nops", C.c_int32), ("accumulate", C.c_int32),
                ("ops", C.POINTER(OpDesc)), ("bc", BcDesc * DEO_MAX_DIMS),
                ("flags", C.c_int32), ("reserved", C.c_int32)]


EXPORTS = {
    # name: (argtypes)
    "deo_abi_version": [],
    "deo_device_count": [C.POINTER(C.c_int32)],
    "deo_init": [C.c_int32],
    "deo_sync": [],
    "deo_last_error": [C.c_char_p, C.c_size_t],
    "deo_launch_count": [C.POINTER(C.c_int64)],
    "deo_buffer_create": [C.c_size_t, C.POINTER(C.c_void_p)],
    "deo_buffer_free": [C.c_void_p],
    "deo_buffer_size": [C.c_void_p, C.POINTER(C.c_size_t)],
    "deo_buffer_upload": [C.c_void_p, C.c_void_p, C.c_size_t],
    "deo_buffer_download": [C.c_void_p, C.c_void_p, C.c_size_t],
    "deo_buffer_devptr": [C.c_void_p, C.POINTER(C.c_void_p)],
    "deo_buffer_wrap": [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)],
    "deo_buffer_muladd": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_int64, C.c_int32],
    "deo_host_alloc": [C.c_size_t, C.POINTER(C.c_void_p)],
    "deo_host_free": [C.c_void_p],
    "deo_plan_create": [C.POINTER(PlanDesc), C.POINTER(C.c_void_p)],
    "deo_plan_destroy": [C.c_void_p],
    "deo_plan_update_coefficients": [C.c_void_p, C.c_int32, C.c_void_p],
    "deo_plan_apply": [C.c_void_p, C.c_void_p, C.c_void_p],
    "deo_plan_apply_axpy": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double],
    "deo_plan_apply_n": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32],
    "deo_plan_apply_host": [C.c_void_p, C.c_void_p, C.c_void_p],
    "deo_plan_info": [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_int32)],
    "deo_plan_time": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_float)],
    "deo_dist_unique_id": [C.c_void_p],
    "deo_dist_init": [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)],
    "deo_dist_destroy": [C.c_void_p],
    "deo_dist_slab": [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
    "deo_dist_plan_create": [C.c_void_p, C.POINTER(PlanDesc), C.POINTER(C.c_void_p)],
    "deo_dist_plan_halo": [C.c_void_p, C.POINTER(C.c_int32)],
    "deo_dist_plan_apply": [C.c_void_p, C.c_void_p, C.c_void_p],
    "deo_dist_plan_apply_host": [C.c_void_p, C.c_void_p, C.c_void_p],
    "deo_dist_plan_time": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_float)],
    "deo_dist_plan_create_local": [C.POINTER(PlanDesc), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)],
}

_lib = None


def load():
    """Load libdeo_b200.so and bind every symbol include/deo_b200.h declares."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(the CUDA library is the only compute path; there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, argtypes in EXPORTS.items():
            f = getattr(L, name)          # AttributeError if a declared symbol is not exported
            f.argtypes = argtypes
            f.restype = C.c_int32
        _lib = L
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(1024)
    load().deo_last_error(buf, len(buf))
    return buf.value.decode(errors="replace")


def check(rc: int):
    if rc != DEO_OK:
        raise DeoError(rc, last_error())


def dtype_code(dtype) -> int:
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return DEO_F64
    if dtype == np.float32:
        return DEO_F32
    raise TypeError(f"the device path requires Float32 or Float64, got {dtype}")


def launch_count() -> int:
    n = C.c_int64(0)
    check(load().deo_launch_count(C.byref(n)))
    return int(n.value)
