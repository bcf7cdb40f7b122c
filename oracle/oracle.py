"""ctypes bindings of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module
(see oracle/ttv_oracle.h).  Nothing under ttv_b200/ imports it.

  Oracle      oracle/libttv_oracle.so         plain-C restatement of the reference algorithm ("port")
  Reference   oracle/_ref/libttv_ref*.so      the unmodified reference headers compiled here ("reference");
                                              prebuilt in the container, shipped to the GPU box by gpurun
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

DTYPES = {
    "f32": (0, np.float32),
    "f64": (1, np.float64),
    "c64": (2, np.complex64),
    "c128": (3, np.complex128),
    "i32": (4, np.int32),
    "i64": (5, np.int64),
}
NP2CODE = {np.dtype(v[1]): v[0] for v in DTYPES.values()}

# numbering shared with include/ttv_b200.h
EXEC = {"seq": 0, "seq_blas": 1, "par": 2, "par_loop": 3, "par_taskloop": 4, "par_task": 5, "par_blas": 6}
SLICING = {"slice": 0, "subtensor": 1}
FUSION = {"none": 0, "outer": 1, "all": 2}

# the 17 combinations reachable through the reference's public wrapper (tensor_times_vector.h:430-1361)
REF_COMBOS = [
    ("seq", "slice", "none"), ("seq_blas", "slice", "none"), ("par_task", "slice", "none"),
    ("par_taskloop", "slice", "none"), ("par", "slice", "none"), ("par_loop", "slice", "none"),
    ("par_loop", "slice", "outer"), ("par_loop", "slice", "all"), ("par_blas", "slice", "all"),
    ("seq", "subtensor", "none"), ("seq_blas", "subtensor", "none"), ("par_task", "subtensor", "none"),
    ("par_taskloop", "subtensor", "none"), ("par", "subtensor", "none"), ("par_loop", "subtensor", "none"),
    ("par_loop", "subtensor", "all"), ("par_blas", "subtensor", "all"),
]

_u64p = C.POINTER(C.c_uint64)


def u64(seq):
    return np.ascontiguousarray(np.asarray(list(seq), dtype=np.uint64))


def _p(arr):
    return arr.ctypes.data_as(_u64p) if arr is not None else None


def build(force: bool = False) -> None:
    """Compile libttv_oracle.so (and oracle/_ref when /root/reference is present). Building the checker is not using it."""
    so = os.path.join(HERE, "libttv_oracle.so")
    src_newer = (not os.path.exists(so)) or any(
        os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(so)
        for f in ("ttv_oracle.c", "ttv_oracle_impl.inc", "ttv_oracle.h"))
    if force or src_newer:
        subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if os.path.isdir("/root/reference/include/tlib"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


class Oracle:
    """The plain-C restatement (sequential).  All arrays are numpy; modes and layouts are 1-based."""

    def __init__(self):
        build()
        self.lib = C.CDLL(os.path.join(HERE, "libttv_oracle.so"))
        L = self.lib
        L.ttv_oracle_run.restype = C.c_int
        L.ttv_oracle_run.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, _u64p, _u64p, _u64p,
                                     C.c_void_p, _u64p, C.c_void_p, _u64p, _u64p, _u64p]
        L.ttv_oracle_naive.restype = C.c_int
        L.ttv_oracle_naive.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, _u64p, _u64p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]
        for name in ("ttv_oracle_gemv_row", "ttv_oracle_gemv_col"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]
        L.ttv_oracle_strerror.restype = C.c_char_p
        L.ttv_oracle_strerror.argtypes = [C.c_int]
        L.ttv_oracle_fill.restype = None
        L.ttv_oracle_fill.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]
        L.ttv_oracle_case.restype = C.c_int
        L.ttv_oracle_case.argtypes = [C.c_uint64, C.c_uint64, _u64p]

    # ---- L0 helpers -------------------------------------------------------------------------------------------
    def is_valid_shape(self, n):
        n = u64(n); return bool(self.lib.ttv_oracle_is_valid_shape(_p(n), C.c_uint64(len(n))))

    def is_valid_layout(self, pi):
        pi = u64(pi); return bool(self.lib.ttv_oracle_is_valid_layout(_p(pi), C.c_uint64(len(pi))))

    def is_valid_strides(self, pi, w):
        pi, w = u64(pi), u64(w)
        return bool(self.lib.ttv_oracle_is_valid_strides(_p(pi), C.c_uint64(len(pi)), _p(w)))

    def strides(self, n, pi):
        n, pi = u64(n), u64(pi); w = np.zeros(len(n), np.uint64)
        if self.lib.ttv_oracle_compute_strides(_p(n), _p(pi), C.c_uint64(len(n)), _p(w)): raise ValueError("invalid shape/layout")
        return [int(x) for x in w]

    def output_shape(self, na, q):
        na = u64(na); nc = np.zeros(max(len(na) - 1, 1), np.uint64)
        if self.lib.ttv_oracle_output_shape(_p(na), C.c_uint64(len(na)), C.c_uint64(q), _p(nc)): raise ValueError("invalid")
        return [int(x) for x in nc[: len(na) - 1]]

    def output_layout(self, pia, q):
        pia = u64(pia); pic = np.zeros(max(len(pia) - 1, 1), np.uint64)
        if self.lib.ttv_oracle_output_layout(_p(pia), C.c_uint64(len(pia)), C.c_uint64(q), _p(pic)): raise ValueError("invalid")
        return [int(x) for x in pic[: len(pia) - 1]]

    def k_order_layout(self, p, k):
        pi = np.zeros(p, np.uint64)
        if self.lib.ttv_oracle_k_order_layout(C.c_uint64(p), C.c_uint64(k), _p(pi)): raise ValueError("invalid")
        return [int(x) for x in pi]

    def case(self, p, q, pia):
        pia = u64(pia); return int(self.lib.ttv_oracle_case(C.c_uint64(p), C.c_uint64(q), _p(pia)))

    def strerror(self, st):
        return self.lib.ttv_oracle_strerror(st).decode()

    # ---- the path -----------------------------------------------------------------------------------------------
    def run_raw(self, dtype, slicing, q, p, a, na, wa, pia, b, nb, c, nc, wc, pic):
        """Thin call with explicit (possibly invalid / None) arguments; returns the status code."""
        def vp(x):
            return x.ctypes.data_as(C.c_void_p) if x is not None else None
        return int(self.lib.ttv_oracle_run(dtype, slicing, q, p, vp(a), _p(na), _p(wa), _p(pia), vp(b), _p(nb),
                                           vp(c), _p(nc), _p(wc), _p(pic)))

    def ttv(self, q, a, na, pia, b, slicing="subtensor"):
        """C = A x_q b for a packed tensor given as flat array `a` with shape na and layout pia. Returns flat C
        in the output layout (packed strides of (nc, pic))."""
        a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
        code = NP2CODE[a.dtype]
        p = len(na)
        nc, pic = self.output_shape(na, q), self.output_layout(pia, q)
        wa, wc = self.strides(na, pia), self.strides(nc, pic)
        c = np.zeros(int(np.prod([int(x) for x in nc], dtype=object)), a.dtype)
        st = self.run_raw(code, SLICING[slicing], q, p, a, u64(na), u64(wa), u64(pia), b, u64([len(b)]),
                          c, u64(nc), u64(wc), u64(pic))
        if st: raise RuntimeError(self.strerror(st))
        return c

    def naive(self, q, a, na, pia, b, want_abs=False):
        a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
        code = NP2CODE[a.dtype]
        n_out = int(np.prod([int(x) for x in na], dtype=object)) // int(na[q - 1])
        c = np.zeros(n_out, a.dtype)
        mag = np.zeros(n_out, np.float64) if want_abs else None
        nav, piav = u64(na), u64(pia)
        st = self.lib.ttv_oracle_naive(code, q, len(na), a.ctypes.data_as(C.c_void_p), _p(nav), _p(piav),
                                       b.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p),
                                       mag.ctypes.data_as(C.c_void_p) if want_abs else None)
        if st: raise RuntimeError("ttv_oracle_naive: bad arguments")
        return (c, mag) if want_abs else c

    def gemv(self, kind, a, b, M, N, lda):
        a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
        c = np.zeros(M, a.dtype)
        f = self.lib.ttv_oracle_gemv_row if kind == "row" else self.lib.ttv_oracle_gemv_col
        if f(NP2CODE[a.dtype], a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p),
             M, N, lda): raise RuntimeError("bad dtype")
        return c

    def fill(self, dtype_name, count, seed, first=0):
        code, npdt = DTYPES[dtype_name]
        x = np.empty(count, npdt)
        self.lib.ttv_oracle_fill(code, x.ctypes.data_as(C.c_void_p), first, count, seed)
        return x


class Reference:
    """The unmodified reference, compiled from /root/reference/include into oracle/_ref (see oracle/Makefile)."""

    def __init__(self, blas: bool = False):
        name = "libttv_ref_openblas.so" if blas else "libttv_ref.so"
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        L = self.lib
        L.ttv_ref_run.restype = C.c_int
        L.ttv_ref_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, _u64p, _u64p,
                                  _u64p, C.c_void_p, _u64p, C.c_void_p, _u64p, _u64p, _u64p]
        L.ttv_ref_tensor.restype = C.c_int
        L.ttv_ref_tensor.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, _u64p, _u64p, C.c_void_p,
                                     C.c_void_p, _u64p, _u64p, _u64p]
        L.ttv_ref_last_error.restype = C.c_char_p
        L.ttv_ref_cores.restype = C.c_uint
        self.blas = bool(L.ttv_ref_has_blas())

    @staticmethod
    def available(blas: bool = False) -> bool:
        return os.path.exists(os.path.join(HERE, "_ref", "libttv_ref_openblas.so" if blas else "libttv_ref.so"))

    def cores(self):
        return int(self.lib.ttv_ref_cores())

    def last_error(self):
        return self.lib.ttv_ref_last_error().decode()

    def run_raw(self, dtype, combo, q, p, a, na, wa, pia, b, nb, c, nc, wc, pic):
        def vp(x):
            return x.ctypes.data_as(C.c_void_p) if x is not None else None
        ep, sp, fp = combo
        return int(self.lib.ttv_ref_run(dtype, EXEC[ep], SLICING[sp], FUSION[fp], q, p, vp(a), _p(na), _p(wa), _p(pia),
                                        vp(b), _p(nb), vp(c), _p(nc), _p(wc), _p(pic)))

    def ttv(self, q, a, na, pia, b, combo=("seq", "subtensor", "none"), helpers: Oracle | None = None, c0=None):
        h = helpers or _default_oracle()
        a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
        nc, pic = h.output_shape(na, q), h.output_layout(pia, q)
        wa, wc = h.strides(na, pia), h.strides(nc, pic)
        c = np.zeros(int(np.prod([int(x) for x in nc], dtype=object)), a.dtype) if c0 is None else c0
        st = self.run_raw(NP2CODE[a.dtype], combo, q, len(na), a, u64(na), u64(wa), u64(pia), b, u64([len(b)]),
                          c, u64(nc), u64(wc), u64(pic))
        if st: raise RuntimeError(self.last_error())
        return c

    def tensor_iface(self, q, a, na, pia, b, use_operator=True):
        a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
        p = len(na)
        c = np.zeros(a.size // int(na[q - 1]), a.dtype)
        nc = np.zeros(p, np.uint64); pic = np.zeros(p, np.uint64); wc = np.zeros(p, np.uint64)
        nav, piav = u64(na), u64(pia)
        st = self.lib.ttv_ref_tensor(NP2CODE[a.dtype], int(use_operator), q, p, a.ctypes.data_as(C.c_void_p), _p(nav),
                                     _p(piav), b.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p),
                                     _p(nc), _p(pic), _p(wc))
        if st: raise RuntimeError(self.last_error())
        return c, [int(x) for x in nc[:p - 1]], [int(x) for x in pic[:p - 1]], [int(x) for x in wc[:p - 1]]


_ORACLE = None


def _default_oracle() -> Oracle:
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE
