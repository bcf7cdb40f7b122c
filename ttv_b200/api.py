"""Python host side of the TTV path on top of the C-ABI (include/ttv_b200.h).

Mirrors the reference's interfaces for this path:

  ttv_lowlevel(...)   the C-like interface   tlib::ttv::ttv(ep, sp, fp, q, p, a, na, wa, pia, b, nb, c, nc, wc, pic)
                      (reference include/tlib/ttv.h:54-92): flat buffers + shape / stride / layout tuples, 1-based modes
  ttv(q, A, b)        the tensor-level interface (ttv.h:99-114) on numpy arrays or torch CUDA tensors; the output
                      shape / layout follow detail/shape.h:126-158 and detail/layout.h:175-207

numpy arrays are HOST buffers (staged through the device inside the call); torch CUDA tensors are used in place.
torch is only used for device memory and streams -- all arithmetic happens in libttv_b200.so.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import functools
from typing import Sequence

import numpy as np

from . import _lib
from ._lib import Opts, Plan

DTYPE_CODES = {"f32": 0, "f64": 1, "c64": 2, "c128": 3, "i32": 4, "i64": 5}
_NP_CODES = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.complex64): 2,
             np.dtype(np.complex128): 3, np.dtype(np.int32): 4, np.dtype(np.int64): 5}

EXECUTION = {"seq": 0, "seq_blas": 1, "par": 2, "par_loop": 3, "par_taskloop": 4, "par_task": 5, "par_blas": 6,
             "par_blas_loop": 7}
SLICING = {"slice": 0, "subtensor": 1}
FUSION = {"none": 0, "outer": 1, "all": 2}
KERNELS = {"auto": 0, "dot": 1, "col": 2, "stream": 3, "colx": 4, "dotf": 5, "colt": 7, "streamk": 8, "dotp": 9, "colf": 10}

FLAG_ACCUMULATE, FLAG_ASYNC, FLAG_NO_VEC = 1, 2, 4


class TTVError(RuntimeError):
    """Raised for every non-zero status of the C-ABI; str() is the reference's message text (ttv.h:64-89)."""

    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


_TORCH_CODES: dict = {}


def dtype_code(x) -> int:
    if _is_torch(x):
        table = _TORCH_CODES
        if not table:
            import torch
            table.update({torch.float32: 0, torch.float64: 1, torch.complex64: 2, torch.complex128: 3, torch.int32: 4,
                          torch.int64: 5})
        if x.dtype not in table:
            raise TTVError(31, f"Error in ttv_b200: unsupported element type {x.dtype}.")
        return table[x.dtype]
    dt = np.asarray(x).dtype
    if dt not in _NP_CODES:
        raise TTVError(31, f"Error in ttv_b200: unsupported element type {dt}.")
    return _NP_CODES[dt]


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, int):
        return C.c_void_p(x)
    if _is_torch(x):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(x.ctypes.data)


@functools.lru_cache(maxsize=8192)
def _tuple_cached(key: tuple):
    arr = np.ascontiguousarray(np.asarray(key, dtype=np.uint64))
    if arr.size == 0:
        arr = np.zeros(1, np.uint64)        # keep the pointer non-null; the length travels separately (p)
    arr.setflags(write=False)
    return arr, C.cast(arr.ctypes.data, _lib.u64p)


def _tuple(v):
    """(uint64 array, pointer) of a shape / stride / layout tuple.  The C side only reads them, so the conversions are
    cached per value: a call of the low-level interface converts seven tuples, which otherwise costs more host time than
    the launch itself."""
    if v is None:
        return None, None
    return _tuple_cached(tuple(int(x) for x in v))


def make_opts(*, execution="par_loop", slicing="subtensor", fusion="all", kernel="auto", ksplit=0, flags=0, device=-1,
              stream=None) -> Opts:
    if stream is not None and not isinstance(stream, int):
        stream = stream.cuda_stream          # torch.cuda.Stream
    return Opts(device=device, execution=EXECUTION[execution] if isinstance(execution, str) else int(execution),
                slicing=SLICING[slicing] if isinstance(slicing, str) else int(slicing),
                fusion=FUSION[fusion] if isinstance(fusion, str) else int(fusion),
                kernel=KERNELS[kernel] if isinstance(kernel, str) else int(kernel),
                ksplit=int(ksplit), flags=int(flags), reserved=0, stream=stream or None)


def _current_torch_stream(x):
    import torch
    return torch.cuda.current_stream(x.device).cuda_stream


def _check(status: int):
    if status:
        lib = _lib.load()
        msg = lib.ttv_b200_last_error().decode() or lib.ttv_b200_strerror(status).decode()
        raise TTVError(status, msg)


# ---- L0 helpers (reference detail/shape.h, layout.h, strides.h) ----------------------------------------------------
def is_valid_shape(n) -> bool:
    arr, p = _tuple(n)
    return bool(_lib.load().ttv_b200_is_valid_shape(p, len(n)))


def is_valid_layout(pi) -> bool:
    arr, p = _tuple(pi)
    return bool(_lib.load().ttv_b200_is_valid_layout(p, len(pi)))


def is_valid_strides(pi, w) -> bool:
    a1, p1 = _tuple(pi); a2, p2 = _tuple(w)
    r = _lib.load().ttv_b200_is_valid_strides(p1, len(pi), p2)
    if r < 0:
        raise TTVError(16, "Error in tlib::detail::is_valid_strides(): input layout is not valid.")
    return bool(r)


# The tuple helpers are pure functions of small integer tuples and sit on the path of every call of the tensor-level
# interface, so their results are cached (a failing call raises every time: lru_cache does not cache exceptions).
@functools.lru_cache(maxsize=8192)
def _strides_cached(n: tuple, pi: tuple) -> tuple:
    a1, p1 = _tuple(n); a2, p2 = _tuple(pi)
    w = np.zeros(max(len(n), 1), np.uint64)
    if len(n) != len(pi) or _lib.load().ttv_b200_compute_strides(p1, p2, len(n), w.ctypes.data_as(_lib.u64p)):
        raise TTVError(14, "Error in tlib::detail::compute_strides(): input shape or layout is not valid.")
    return tuple(int(x) for x in w[: len(n)])


def generate_strides(n, pi) -> list[int]:
    return list(_strides_cached(tuple(int(x) for x in n), tuple(int(x) for x in pi)))


@functools.lru_cache(maxsize=8192)
def _output_shape_cached(na: tuple, q: int) -> tuple:
    a1, p1 = _tuple(na)
    nc = np.zeros(max(len(na), 1), np.uint64)
    if _lib.load().ttv_b200_output_shape(p1, len(na), q, nc.ctypes.data_as(_lib.u64p)):
        raise TTVError(14, "Error in tlib::detail::generate_output_shape(): input shape or contraction mode is not valid.")
    return tuple(int(x) for x in nc[: len(na) - 1])


def generate_output_shape(na, q) -> list[int]:
    return list(_output_shape_cached(tuple(int(x) for x in na), int(q)))


@functools.lru_cache(maxsize=8192)
def _output_layout_cached(pia: tuple, q: int) -> tuple:
    a1, p1 = _tuple(pia)
    pic = np.zeros(max(len(pia), 1), np.uint64)
    if _lib.load().ttv_b200_output_layout(p1, len(pia), q, pic.ctypes.data_as(_lib.u64p)):
        raise TTVError(16, "Error in tlib::detail::generate_output_layout(): input layout or contraction mode is not valid.")
    return tuple(int(x) for x in pic[: len(pia) - 1])


def generate_output_layout(pia, q) -> list[int]:
    return list(_output_layout_cached(tuple(int(x) for x in pia), int(q)))


@functools.lru_cache(maxsize=1024)
def _k_order_cached(p: int, k: int) -> tuple:
    pi = np.zeros(max(p, 1), np.uint64)
    if _lib.load().ttv_b200_k_order_layout(p, k, pi.ctypes.data_as(_lib.u64p)):
        raise TTVError(16, "Error in tlib::detail::compute_k_order: range provided by begin and end not correct!")
    return tuple(int(x) for x in pi[:p])


def generate_k_order_layout(p, k) -> list[int]:
    return list(_k_order_cached(int(p), int(k)))


def _is_contiguous(x) -> bool:
    return bool(x.is_contiguous()) if _is_torch(x) else bool(np.asarray(x).flags.c_contiguous or np.asarray(x).flags.f_contiguous)


def _check_operands(a, b, c):
    """a, b, c travel as raw pointers: whatever is array-like must agree in element type and in where it lives, and b / c
    must be dense -- otherwise the library would silently read another type's bytes or write past the end of c."""
    named = [(n, x) for n, x in (("a", a), ("b", b), ("c", c)) if x is not None and not isinstance(x, int)]
    if not named:
        return
    codes = {n: dtype_code(x) for n, x in named}
    if len(set(codes.values())) > 1:
        names = {n: (str(x.dtype)) for n, x in named}
        raise TTVError(31, f"Error in ttv_b200: a, b and c must have the same element type, got {names}.")
    kinds = {n: (("cuda:%d" % x.device.index) if (_is_torch(x) and x.is_cuda) else "host") for n, x in named}
    if len(set(kinds.values())) > 1:
        raise TTVError(41, f"Error in ttv_b200: a, b and c must all be host buffers or all live on one device, got {kinds}.")
    for n, x in named:
        if n != "a" and not _is_contiguous(x):
            raise TTVError(32, f"Error in ttv_b200: {n} must be a dense (contiguous) buffer.")


# ---- the low-level interface -----------------------------------------------------------------------------------------
def ttv_lowlevel(q: int, p: int, a, na, wa, pia, b, nb, c, nc, wc, pic, *, dtype: int | None = None,
                 opts: Opts | None = None, **opt_kwargs) -> None:
    """The reference's C-like interface (ttv.h:54-92).  a, b, c: numpy arrays (host), torch CUDA tensors (device),
    raw integer addresses, or None; the tuples are sequences of ints or None.  Raises TTVError with the
    reference's message on invalid arguments.  C is overwritten."""
    lib = _lib.load()
    _check_operands(a, b, c)
    if dtype is None:
        probe = next((x for x in (a, b, c) if x is not None and not isinstance(x, int)), None)
        if probe is None:
            raise ValueError("dtype is required when a, b, c are raw addresses")
        dtype = dtype_code(probe)
    if opts is None:
        if "stream" not in opt_kwargs:
            dev = next((x for x in (a, b, c) if x is not None and _is_torch(x) and x.is_cuda), None)
            if dev is not None:
                opt_kwargs["stream"] = _current_torch_stream(dev)
        opts = make_opts(**opt_kwargs)
    keep = [_tuple(v) for v in (na, wa, pia, nb, nc, wc, pic)]
    (na_, wa_, pia_, nb_, nc_, wc_, pic_) = [k[1] for k in keep]
    st = lib.ttv_b200_run(dtype, q, p, _ptr(a), na_, wa_, pia_, _ptr(b), nb_, _ptr(c), nc_, wc_, pic_, C.byref(opts))
    _check(st)


def prepared_lowlevel(q: int, p: int, a, na, wa, pia, b, nb, c, nc, wc, pic, *, opts: Opts | None = None, **opt_kwargs):
    """ttv_lowlevel with everything that does not change between calls done ONCE (operand checks, tuple conversion, the opts
    block): returns `run()`, which is one ctypes call of ttv_b200_run -- a few microseconds of host time instead of the
    ~40 of the convenience wrapper.  For loops over short products (a 512 MiB tensor is an 80 us kernel) where the host
    must stay ahead of the GPU.  The operands are captured: refill them in place between calls."""
    lib = _lib.load()
    _check_operands(a, b, c)
    dtype = dtype_code(next(x for x in (a, b, c) if x is not None and not isinstance(x, int)))
    if opts is None:
        if "stream" not in opt_kwargs:
            dev = next((x for x in (a, b, c) if x is not None and _is_torch(x) and x.is_cuda), None)
            if dev is not None:
                opt_kwargs["stream"] = _current_torch_stream(dev)
        opts = make_opts(**opt_kwargs)
    keep = [_tuple(v) for v in (na, wa, pia, nb, nc, wc, pic)]
    (na_, wa_, pia_, nb_, nc_, wc_, pic_) = [k[1] for k in keep]
    pa, pb, pc = _ptr(a), _ptr(b), _ptr(c)
    ref = C.byref(opts)
    fn = lib.ttv_b200_run

    def run(_keep=(a, b, c, keep, opts)):
        st = fn(dtype, q, p, pa, na_, wa_, pia_, pb, nb_, pc, nc_, wc_, pic_, ref)
        if st:
            _check(st)

    return run


class Resident:
    """Device-side twin of ONE host tensor (ttv_b200_resident): the first product after creation / invalidate() streams A
    across PCIe under its own kernels into a buffer that stays in HBM; later products of the same array only move b and C.
    This is what tlib::ttv::tensor::keep_on_device(true) uses behind `A(q) * b`; call invalidate() when the host data
    changed."""

    def __init__(self, device: int = -1):
        self._lib = _lib.load()
        h = C.c_void_p()
        _check(self._lib.ttv_b200_resident_create(C.byref(h), int(device)))
        self._h = h

    def invalidate(self) -> None:
        self._lib.ttv_b200_resident_invalidate(self._h)

    @property
    def valid(self) -> bool:
        return bool(self._lib.ttv_b200_resident_valid(self._h))

    def ttv_lowlevel(self, q: int, p: int, a, na, wa, pia, b, nb, c, nc, wc, pic, *, opts: Opts | None = None, **opt_kwargs) -> None:
        """ttv_lowlevel with HOST buffers whose A keeps its copy on the device (ttv_b200_run_resident)"""
        _check_operands(a, b, c)
        if opts is None:
            opts = make_opts(**opt_kwargs)
        keep = [_tuple(v) for v in (na, wa, pia, nb, nc, wc, pic)]
        (na_, wa_, pia_, nb_, nc_, wc_, pic_) = [k[1] for k in keep]
        _check(self._lib.ttv_b200_run_resident(self._h, dtype_code(a), q, p, _ptr(a), na_, wa_, pia_, _ptr(b), nb_, _ptr(c), nc_, wc_,
                                               pic_, C.byref(opts)))

    def close(self) -> None:
        if self._h is not None and self._h.value:
            self._lib.ttv_b200_resident_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ttv_lowlevel_devices(devices: Sequence[int], q: int, p: int, a, na, wa, pia, b, nb, c, nc, wc, pic, *, opts: Opts | None = None,
                         **opt_kwargs) -> None:
    """ttv_lowlevel with HOST buffers, the work cut along the slowest mode of A over several GPUs of this process, one
    host thread and one PCIe link per GPU (ttv_b200_run_devices)."""
    lib = _lib.load()
    _check_operands(a, b, c)
    if opts is None:
        opts = make_opts(**opt_kwargs)
    keep = [_tuple(v) for v in (na, wa, pia, nb, nc, wc, pic)]
    (na_, wa_, pia_, nb_, nc_, wc_, pic_) = [k[1] for k in keep]
    devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
    _check(lib.ttv_b200_run_devices(dtype_code(a), q, p, _ptr(a), na_, wa_, pia_, _ptr(b), nb_, _ptr(c), nc_, wc_, pic_, C.byref(opts),
                                    devs, len(devices)))


def pinned_empty(count: int, dtype) -> np.ndarray:
    """a numpy array in page-locked host memory (ttv_b200_host_alloc): H2D / D2H run by DMA straight from it.  The memory is
    released when the array (and every view of it) is gone."""
    lib = _lib.load()
    dt = np.dtype(dtype)
    nbytes = int(count) * dt.itemsize
    ptr = C.c_void_p()
    _check(lib.ttv_b200_host_alloc(C.byref(ptr), nbytes))

    import weakref
    buf = (C.c_char * max(nbytes, 1)).from_address(ptr.value)
    weakref.finalize(buf, lib.ttv_b200_host_free, C.c_void_p(ptr.value))      # numpy keeps `buf` alive through arr.base
    return np.frombuffer(buf, dtype=dt, count=int(count))


def ttv_multi(qs: Sequence[int], a, na, pia, bs, cs=None, *, wa=None, opts: Opts | None = None, **opt_kwargs):
    """C_i = A x_{qs[i]} bs[i] for several modes of ONE tensor (ttv_b200_multi).  a: flat buffer (numpy = host,
    torch CUDA = device) with shape na and layout pia; bs: vectors of the same kind.  With host buffers A is copied to
    the device once for all products.  Returns the list of flat C_i (packed in the output layout of each q)."""
    lib = _lib.load()
    p = len(na)
    wa = list(wa) if wa is not None else generate_strides(na, pia)
    torch_in = _is_torch(a)
    if cs is None:
        cs = []
        for q in qs:
            n_out = int(np.prod(generate_output_shape(na, q), dtype=object))
            if torch_in:
                import torch
                cs.append(torch.empty(n_out, dtype=a.dtype, device=a.device))
            else:
                cs.append(np.empty(n_out, dtype=np.asarray(a).dtype))
    if opts is None:
        if "stream" not in opt_kwargs and torch_in and a.is_cuda:
            opt_kwargs["stream"] = _current_torch_stream(a)
        opts = make_opts(**opt_kwargs)
    keep = [_tuple(v) for v in (na, wa, pia, list(qs))]
    n = len(qs)
    barr = (C.c_void_p * n)(*[_ptr(b).value for b in bs])
    carr = (C.c_void_p * n)(*[_ptr(c).value for c in cs])
    _check(lib.ttv_b200_multi(dtype_code(a), p, _ptr(a), keep[0][1], keep[1][1], keep[2][1], n, keep[3][1], barr, carr,
                              C.byref(opts)))
    return cs


CHAIN_ORDERS = {"optimal": 0, "backward": 1, "forward": 2}


def chain_plan(q: int, na, order: str = "optimal"):
    """[(mode to contract, index into bs), ...] of the p-1 products ttv_b200_ttvs runs (ttv_b200_chain_plan; pure host
    code).  Modes are numbered in the tensor that is left when the step runs (wrapped_ttv.cpp:135-192)."""
    lib = _lib.load()
    p = len(na)
    modes = (C.c_uint64 * max(p - 1, 1))()
    vecs = (C.c_uint64 * max(p - 1, 1))()
    _check(lib.ttv_b200_chain_plan(q, p, _tuple(na)[1], CHAIN_ORDERS[order], modes, vecs))
    return [(int(modes[i]), int(vecs[i])) for i in range(p - 1)]


def ttvs(q: int, a, na, pia, bs, order: str = "optimal", out=None, *, opts: Opts | None = None, **opt_kwargs):
    """c = A contracted with bs[j] along every mode except q (ttv_b200_ttvs, the native form of ttvpy::ttvs,
    wrapped_ttv.cpp:83-198).  a: flat packed buffer of shape na / layout pia -- numpy (host: A crosses PCIe once, the
    intermediates stay in HBM) or torch CUDA (device).  bs: the p-1 vectors in mode order, same kind as a.  Returns the
    vector of na[q-1] elements (`out` if given)."""
    lib = _lib.load()
    p = len(na)
    torch_in = _is_torch(a)
    if out is None:
        n_out = int(na[q - 1]) if 1 <= q <= p else 1        # (an invalid q is reported by the library, with its text)
        if torch_in:
            import torch
            out = torch.empty(n_out, dtype=a.dtype, device=a.device)
        else:
            out = np.empty(n_out, dtype=np.asarray(a).dtype)
    if opts is None:
        if "stream" not in opt_kwargs and torch_in and a.is_cuda:
            opt_kwargs["stream"] = _current_torch_stream(a)
        opts = make_opts(**opt_kwargs)
    n = len(bs)
    barr = (C.c_void_p * max(n, 1))(*[_ptr(b).value for b in bs])
    _check(lib.ttv_b200_ttvs(dtype_code(a), q, p, _ptr(a), _tuple(na)[1], _tuple(pia)[1], barr, CHAIN_ORDERS[order], _ptr(out),
                             C.byref(opts)))
    return out


def plan(q: int, na, pia, *, dtype="f32", wa=None, wc=None, pic=None, **opt_kwargs) -> dict:
    """What the layout folder and the kernel chooser decide for (na, pia, q): pure host code, needs no GPU."""
    lib = _lib.load()
    p = len(na)
    code = DTYPE_CODES[dtype] if isinstance(dtype, str) else int(dtype)
    nc = generate_output_shape(na, q) if p > 1 else [1]
    pic = list(pic) if pic is not None else (generate_output_layout(pia, q) if p > 1 else [1])
    wa = list(wa) if wa is not None else generate_strides(na, pia)
    wc = list(wc) if wc is not None else (generate_strides(nc, pic) if p > 1 else [1])
    keep = [_tuple(v) for v in (na, wa, pia, [na[q - 1]], nc, wc, pic)]
    (na_, wa_, pia_, nb_, nc_, wc_, pic_) = [k[1] for k in keep]
    out = Plan()
    opts = make_opts(**opt_kwargs)
    one = C.c_void_p(16)   # any non-null value: plan never dereferences a, b, c
    _check(lib.ttv_b200_plan(code, q, p, one, na_, wa_, pia_, one, nb_, one, nc_, wc_, pic_, C.byref(opts), C.byref(out)))
    return out.as_dict()


def plan_view(outer: int, nq: int, inner: int, *, dtype="f32", **opt_kwargs) -> dict:
    lib = _lib.load()
    code = DTYPE_CODES[dtype] if isinstance(dtype, str) else int(dtype)
    out = Plan()
    opts = make_opts(**opt_kwargs)
    _check(lib.ttv_b200_plan_view(code, outer, nq, inner, C.byref(opts), C.byref(out)))
    return out.as_dict()


def ttv_view(outer: int, nq: int, inner: int, a, b, c, **opt_kwargs) -> None:
    """C[outer][inner] = sum_k A[outer][k][inner] b[k] on DEVICE tensors (the canonical view every legal input folds to)."""
    lib = _lib.load()
    if "stream" not in opt_kwargs and _is_torch(a):
        opt_kwargs["stream"] = _current_torch_stream(a)
    opts = make_opts(**opt_kwargs)
    _check(lib.ttv_b200_view(dtype_code(a), outer, nq, inner, _ptr(a), _ptr(b), _ptr(c), C.byref(opts)))


def ttv_view_scatter(outer: int, nq: int, inner: int, a, b, peer_ptrs, rank: int, blk: int, **opt_kwargs) -> None:
    """This GPU's partial of C[outer][inner] written block by block into the peers' workspaces (ttv_b200_view_scatter):
    peer_ptrs[j] = address of GPU j's workspace [world][blk] as seen from this process.  Asynchronous on the current
    torch stream (a barrier across the GPUs has to follow anyway)."""
    lib = _lib.load()
    if "stream" not in opt_kwargs and _is_torch(a):
        opt_kwargs["stream"] = _current_torch_stream(a)
    opt_kwargs["flags"] = int(opt_kwargs.get("flags", 0)) | 2
    opts = make_opts(**opt_kwargs)
    world = len(peer_ptrs)
    arr = (C.c_void_p * world)(*[int(p) for p in peer_ptrs])
    _check(lib.ttv_b200_view_scatter(dtype_code(a), outer, nq, inner, _ptr(a), _ptr(b), arr, world, rank, blk, C.byref(opts)))


def ttv_view_exchange(outer: int, nq: int, inner: int, a, b, peer_ptrs, flag_ptrs, rank: int, blk: int, c_block, token: int, scratch,
                      max_ctas: int = 0, **opt_kwargs) -> None:
    """The whole n_q-split exchange in ONE kernel per GPU (ttv_b200_view_exchange): product + scatter into the peers' slots,
    in-kernel flag barrier across the GPUs, sum of the received slots into c_block.  Asynchronous on the current torch stream."""
    lib = _lib.load()
    if "stream" not in opt_kwargs and _is_torch(a):
        opt_kwargs["stream"] = _current_torch_stream(a)
    opt_kwargs["flags"] = int(opt_kwargs.get("flags", 0)) | 2
    opts = make_opts(**opt_kwargs)
    world = len(peer_ptrs)
    ws = (C.c_void_p * world)(*[int(p) for p in peer_ptrs])
    fl = (C.c_void_p * world)(*[int(p) for p in flag_ptrs])
    n_block = int(c_block.numel()) if c_block is not None else 0
    _check(lib.ttv_b200_view_exchange(dtype_code(a), outer, nq, inner, _ptr(a), _ptr(b), ws, fl, world, rank, blk, _ptr(c_block), n_block,
                                      int(token), _ptr(scratch), int(max_ctas), C.byref(opts)))


def reduce_slots(ws, c, n: int, blk: int, slots: int, **opt_kwargs) -> None:
    """c[j] = sum over the `slots` rows of ws [slots][blk], j < n, in row order (ttv_b200_reduce_slots); asynchronous on the
    current torch stream."""
    lib = _lib.load()
    if "stream" not in opt_kwargs and _is_torch(c):
        opt_kwargs["stream"] = _current_torch_stream(c)
    opt_kwargs["flags"] = int(opt_kwargs.get("flags", 0)) | 2
    opts = make_opts(**opt_kwargs)
    _check(lib.ttv_b200_reduce_slots(dtype_code(c), _ptr(ws), _ptr(c), n, blk, slots, C.byref(opts)))


def fill(x, seed: int, first: int = 0, count: int | None = None) -> None:
    """x[i] = synth(seed, first + i) on the device (torch CUDA tensor, flat)."""
    lib = _lib.load()
    n = x.numel() if count is None else count
    opts = make_opts(stream=_current_torch_stream(x))
    _check(lib.ttv_b200_fill(dtype_code(x), _ptr(x), first, n, seed, C.byref(opts)))


# ---- the tensor-level interface ----------------------------------------------------------------------------------------
def _element_strides(x):
    """strides of a numpy array / torch tensor in elements, or None when they are not whole positive element counts"""
    if _is_torch(x):
        st = [int(v) for v in x.stride()]
    else:
        if any(v % x.itemsize for v in x.strides):
            return None
        st = [int(v) // x.itemsize for v in x.strides]
    if any(v <= 0 and n > 1 for v, n in zip(st, x.shape)):
        return None                                   # reversed or broadcast axes: not a layout
    return st


def _layout_of(x, layout):
    """(pia, wa, honor) for an array.  A C-contiguous array is a last-order tensor (what ttvpy assumes,
    wrapped_ttv.cpp:44-45), an F-contiguous one a first-order tensor; any other array with positive strides that do not
    overlap is a tensor whose layout is the order of its strides, possibly padded (slices, transposes): it is read IN
    PLACE through wa with TTV_B200_FLAG_HONOR_STRIDES.  Returns None when the array has to be copied first."""
    shape = tuple(int(v) for v in x.shape)
    if layout is not None:
        layout = tuple(int(v) for v in layout)
        if len(layout) != len(shape):
            raise TTVError(16, "Error in tlib::tensor: shape vector and layout vector must have the same length.")
        return list(layout), generate_strides(shape, layout), False
    c_contig = x.is_contiguous() if _is_torch(x) else x.flags.c_contiguous
    st = None if c_contig else _element_strides(x)
    if not c_contig and st is None:
        return None
    desc = _layout_core(shape, bool(c_contig), None if st is None else tuple(st))
    return None if desc is None else (list(desc[0]), list(desc[1]), desc[2])


@functools.lru_cache(maxsize=8192)
def _layout_core(shape: tuple, c_contig: bool, st):
    p = len(shape)
    shape = list(shape)
    if c_contig:
        pia = generate_k_order_layout(p, 0)
        return tuple(pia), tuple(generate_strides(shape, pia)), False
    st = list(st)
    # layout = modes by ascending stride (extent-1 modes last: their stride is meaningless)
    order = sorted(range(p), key=lambda m: (shape[m] == 1, st[m], m))
    need = 1
    for m in order:
        if shape[m] > 1:
            if st[m] < need:
                return None                           # overlapping elements
            need = st[m] * shape[m]
    pia = [m + 1 for m in order]
    wa = [max(1, v) for v in st]
    top = 1
    for m in order:                                   # extent-1 modes: give them a stride that keeps wa valid
        if shape[m] == 1:
            wa[m] = top
        top = max(top, wa[m] * shape[m])
    packed = wa == generate_strides(shape, pia) or all(shape[m] == 1 or wa[m] == w for m, w in enumerate(generate_strides(shape, pia)))
    return tuple(pia), tuple(wa), not packed


def ttv(q: int, A, b, *, layout: Sequence[int] | None = None, out=None, **opt_kwargs):
    """C = A x_q b (1-based q).  A: numpy array (host) or torch CUDA tensor of order p >= 2; b: vector of length
    A.shape[q-1] of the same kind.  `layout` is the 1-based layout tuple of A's memory; by default it is derived from
    the array's contiguity.  With an explicit `layout` A must be a FLAT buffer plus `shape=` in opt_kwargs.
    Returns an array of the same kind with shape A.shape minus mode q, stored in the output layout."""
    shape = opt_kwargs.pop("shape", None)
    torch_in = _is_torch(A)
    if not torch_in:
        A = np.asarray(A)
    wa = None
    if shape is None:
        shape = [int(s) for s in A.shape]
        desc = _layout_of(A, layout)
        if desc is None:                               # reversed / broadcast / overlapping axes: pack a copy
            A = A.contiguous() if torch_in else np.ascontiguousarray(A)
            desc = _layout_of(A, None)
        pia, wa, honor = desc
        if honor:
            opt_kwargs["flags"] = int(opt_kwargs.get("flags", 0)) | 8        # TTV_B200_FLAG_HONOR_STRIDES
    else:
        shape = [int(s) for s in shape]
        pia = [int(v) for v in layout] if layout is not None else generate_k_order_layout(len(shape), 1)
    p = len(shape)
    if p == 0:
        raise TTVError(1, _lib.load().ttv_b200_strerror(1).decode())
    if q == 0 or q > p:
        raise TTVError(2, _lib.load().ttv_b200_strerror(2).decode())
    if p == 1:
        raise TTVError(15, _lib.load().ttv_b200_strerror(15).decode())
    nc = generate_output_shape(shape, q)
    pic = generate_output_layout(pia, q)
    if wa is None:
        wa = generate_strides(shape, pia)
    wc = generate_strides(nc, pic)
    n_out = int(np.prod(nc, dtype=object))

    if torch_in:
        import torch
        if not A.is_cuda:
            raise TTVError(40, "Error in ttv_b200: torch tensors must live on a CUDA device (there is no CPU fallback).")
        if not _is_torch(b):
            raise TTVError(41, "Error in ttv_b200: A is a torch tensor, so b must be one too (same device).")
        if b.dim() == 1 and not b.is_contiguous():
            b = b.contiguous()
        flat_c = out if out is not None else torch.empty(n_out, dtype=A.dtype, device=A.device)
    else:
        A = np.asarray(A)
        if _is_torch(b):
            raise TTVError(41, "Error in ttv_b200: A is a host array, so b must be one too.")
        b = np.asarray(b)
        if b.ndim == 1 and not b.flags.c_contiguous:
            b = np.ascontiguousarray(b)
        flat_c = out if out is not None else np.empty(n_out, dtype=A.dtype)
    if b.ndim > 1:
        raise TTVError(13, "Error in ttv_b200: b must be a vector (one-dimensional).")
    nb = [int(b.shape[0])] if b.ndim >= 1 else [1]
    if out is not None:
        n_have = int(out.numel()) if _is_torch(out) else int(np.asarray(out).size)
        if n_have < n_out:
            raise TTVError(15, f"Error in ttv_b200: out holds {n_have} elements, the product has {n_out}.")
    ttv_lowlevel(q, p, A, shape, wa, pia, b, nb, flat_c, nc, wc, pic, **opt_kwargs)
    if out is not None:
        return out
    # hand the flat result back as an array of shape nc whose memory order is the output layout
    order_slow_to_fast = [m - 1 for m in reversed(pic)]          # axes from slowest to fastest
    packed_shape = [nc[ax] for ax in order_slow_to_fast]
    inv = np.argsort(order_slow_to_fast)
    if torch_in:
        return flat_c.view(*packed_shape).permute(*[int(i) for i in inv])
    return flat_c.reshape(packed_shape).transpose([int(i) for i in inv])


def launch_count() -> int:
    return int(_lib.load().ttv_b200_launch_count())


def device_count() -> int:
    return int(_lib.load().ttv_b200_device_count())
