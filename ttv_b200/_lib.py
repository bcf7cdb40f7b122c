"""Loader and ctypes signatures of libttv_b200.so (the C-ABI declared in include/ttv_b200.h).

The library is built in-tree by ttv_b200/build.py (nvcc, sm_100a).  There is no fallback: if it cannot be loaded the
import fails loudly, and every compute entry fails with TTV_B200_ERR_CUDA when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libttv_b200.so")

u64p = C.POINTER(C.c_uint64)


class Opts(C.Structure):
    """struct ttv_b200_opts"""
    _fields_ = [("device", C.c_int32), ("execution", C.c_int32), ("slicing", C.c_int32), ("fusion", C.c_int32),
                ("kernel", C.c_int32), ("ksplit", C.c_int32), ("flags", C.c_uint32), ("reserved", C.c_int32),
                ("stream", C.c_void_p)]


class Plan(C.Structure):
    """struct ttv_b200_plan_t"""
    _fields_ = [("outer", C.c_uint64), ("nq", C.c_uint64), ("inner", C.c_uint64), ("k", C.c_uint32),
                ("ref_case", C.c_uint32), ("kernel", C.c_int32), ("vec", C.c_int32), ("tx", C.c_int32),
                ("ty", C.c_int32), ("to", C.c_int32), ("nu", C.c_int32), ("ku", C.c_int32), ("ksplit", C.c_int32),
                ("threads", C.c_int32), ("stream", C.c_int32), ("ctas", C.c_uint64),
                ("smem_bytes", C.c_uint64), ("algo_bytes", C.c_uint64), ("algo_flops", C.c_uint64),
                ("workspace_bytes", C.c_uint64)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


# every symbol include/ttv_b200.h declares: name -> (restype, argtypes)
_RUN_ARGS = [C.c_uint64, C.c_uint64, C.c_void_p, u64p, u64p, u64p, C.c_void_p, u64p, C.c_void_p, u64p, u64p, u64p,
             C.POINTER(Opts)]
SYMBOLS = {
    "ttv_b200_run": (C.c_int, [C.c_int] + _RUN_ARGS),
    "ttv_b200_f32": (C.c_int, _RUN_ARGS),
    "ttv_b200_f64": (C.c_int, _RUN_ARGS),
    "ttv_b200_c64": (C.c_int, _RUN_ARGS),
    "ttv_b200_c128": (C.c_int, _RUN_ARGS),
    "ttv_b200_i32": (C.c_int, _RUN_ARGS),
    "ttv_b200_i64": (C.c_int, _RUN_ARGS),
    "ttv_b200_multi": (C.c_int, [C.c_int, C.c_uint64, C.c_void_p, u64p, u64p, u64p, C.c_uint64, u64p,
                                 C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(Opts)]),
    "ttv_b200_ttvs": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, u64p, u64p, C.POINTER(C.c_void_p), C.c_int,
                                C.c_void_p, C.POINTER(Opts)]),
    "ttv_b200_chain_plan": (C.c_int, [C.c_uint64, C.c_uint64, u64p, C.c_int, u64p, u64p]),
    "ttv_b200_plan": (C.c_int, [C.c_int] + _RUN_ARGS + [C.POINTER(Plan)]),
    "ttv_b200_view": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.POINTER(Opts)]),
    "ttv_b200_plan_view": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(Opts), C.POINTER(Plan)]),
    "ttv_b200_view_scatter": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p),
                                        C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(Opts)]),
    "ttv_b200_view_exchange": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p),
                                         C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32,
                                         C.c_void_p, C.c_uint32, C.POINTER(Opts)]),
    "ttv_b200_reduce_slots": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(Opts)]),
    "ttv_b200_fill": (C.c_int, [C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(Opts)]),
    "ttv_b200_device_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint64, C.c_int, C.c_int]),
    "ttv_b200_device_free": (C.c_int, [C.c_void_p]),
    "ttv_b200_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint64]),
    "ttv_b200_host_free": (C.c_int, [C.c_void_p]),
    "ttv_b200_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(Opts)]),
    "ttv_b200_resident_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "ttv_b200_resident_destroy": (None, [C.c_void_p]),
    "ttv_b200_resident_invalidate": (None, [C.c_void_p]),
    "ttv_b200_resident_valid": (C.c_int, [C.c_void_p]),
    "ttv_b200_run_resident": (C.c_int, [C.c_void_p, C.c_int] + _RUN_ARGS),
    "ttv_b200_run_devices": (C.c_int, [C.c_int] + _RUN_ARGS + [C.POINTER(C.c_int32), C.c_uint32]),
    "ttv_b200_is_valid_shape": (C.c_int, [u64p, C.c_uint64]),
    "ttv_b200_is_valid_layout": (C.c_int, [u64p, C.c_uint64]),
    "ttv_b200_is_valid_strides": (C.c_int, [u64p, C.c_uint64, u64p]),
    "ttv_b200_compute_strides": (C.c_int, [u64p, u64p, C.c_uint64, u64p]),
    "ttv_b200_output_shape": (C.c_int, [u64p, C.c_uint64, C.c_uint64, u64p]),
    "ttv_b200_output_layout": (C.c_int, [u64p, C.c_uint64, C.c_uint64, u64p]),
    "ttv_b200_k_order_layout": (C.c_int, [C.c_uint64, C.c_uint64, u64p]),
    "ttv_b200_strerror": (C.c_char_p, [C.c_int]),
    "ttv_b200_last_error": (C.c_char_p, []),
    "ttv_b200_version": (C.c_int, []),
    "ttv_b200_device_count": (C.c_int, []),
    "ttv_b200_launch_count": (C.c_uint64, []),
    "ttv_b200_dtype_size": (C.c_int, [C.c_int]),
    "ttv_b200_release": (None, []),
}

_lib = None


def load() -> C.CDLL:
    """Loads libttv_b200.so; builds it first when nvcc and the sources are there and the library is missing/stale."""
    global _lib
    if _lib is not None:
        return _lib
    override = os.environ.get("TTV_B200_LIB")       # experiments: an alternative build of the same library
    if override:
        lib = C.CDLL(override)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
        return lib
    from . import build as _build
    _build.build_or_warn()        # no nvcc on this box: fine if a prebuilt library travelled with the tree (warns if stale)
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch; fail loudly
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib
