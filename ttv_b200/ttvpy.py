"""Drop-in for the reference's Python binding `ttvpy` (reference ttvpy/src/wrapped_ttv.cpp), on the C-ABI.

    ttv(q, A, b)                      one mode-q product                    wrapped_ttv.cpp:18-78
    ttvs(q, A, bs, order="optimal")   the chain of p-1 products that leaves mode q; order in
                                      {"optimal", "backward", "forward"}      wrapped_ttv.cpp:83-198

Like the reference, a C-contiguous A is a last-order tensor; unlike the reference (float64 only,
wrapped_ttv.cpp:205-206; strides ignored, :44-45) every element type of the C-ABI is accepted and arrays that are not
C-contiguous (transposes, slices, Fortran order) are read in place through their strides.  Errors the reference raises as
std::invalid_argument surface as ValueError with the same text.

ttvs keeps the intermediates in HBM (ttv_b200_ttvs, one native call): A crosses PCIe once (streamed in chunks under the
first product when it comes from host memory), the remaining kernels run back to back on the device, and only the final
vector is copied back.
"""
from __future__ import annotations

import os

import numpy as np

from . import api

__all__ = ["ttv", "ttvs", "chain_plan", "CapturedTtvs"]


def _common_dtype(*arrays):
    """The element type a product of these operands is computed in: numpy's promotion of ALL operands, mapped to an
    element type of the C-ABI.  The reference binds `py::array_t<double>` (wrapped_ttv.cpp:205-206) and lets pybind11 convert
    every operand to float64; promoting keeps that meaning (an integer A with a fractional or complex b is NOT truncated to
    A's type) while float32 / complex / integer operands that agree keep their own type (an extension: the reference would
    hand back the same values as float64)."""
    dt = np.result_type(*[np.asarray(x).dtype if not isinstance(x, np.dtype) else x for x in arrays])
    if dt in api._NP_CODES:
        return dt
    if dt.kind == "c":
        return np.dtype(np.complex128)
    if dt.kind == "f":
        return np.dtype(np.float32) if dt.itemsize < 4 else np.dtype(np.float64)
    if dt.kind in "iub":
        if dt.kind != "u" and dt.itemsize <= 4 or dt.kind == "b" or (dt.kind == "u" and dt.itemsize < 4):
            return np.dtype(np.int32)
        if dt.kind == "i" or dt.itemsize < 8:
            return np.dtype(np.int64)
    return np.dtype(np.float64)             # uint64, objects that convert, ...: what the reference does with everything


def _as_c_array(x, dtype=None):
    a = np.asarray(x)
    return np.ascontiguousarray(a, dtype=dtype if dtype is not None else _common_dtype(a))


def _as_array(x, dtype=None):
    """like _as_c_array but WITHOUT packing: slices and transposes are read in place through their strides (the
    reference silently reads them as if they were C-contiguous, wrapped_ttv.cpp:44-45)"""
    a = np.asarray(x)
    dtype = dtype if dtype is not None else _common_dtype(a)
    return a if a.dtype == dtype else a.astype(dtype)


def _torch_common(A, bs):
    """torch operands: promote like numpy does (torch.result_type pairwise), on A's device"""
    import torch
    dt = A.dtype
    for b in bs:
        dt = torch.promote_types(dt, b.dtype if api._is_torch(b) else torch.from_numpy(np.asarray(b)).dtype)
    if dt not in (torch.float32, torch.float64, torch.complex64, torch.complex128, torch.int32, torch.int64):
        dt = torch.float64 if not dt.is_complex else torch.complex128
    conv = lambda x: (x if api._is_torch(x) else torch.from_numpy(np.asarray(x))).to(device=A.device, dtype=dt)
    return conv(A), [conv(b).contiguous() for b in bs]


def ttv(q: int, A, b):
    """Tensor-times-vector for the q-th mode (1-based) of a numpy array (host) or a torch CUDA tensor (device)."""
    if api._is_torch(A):
        p = A.dim()
        if p == 0:
            raise ValueError("Error calling ttvpy::ttv: input tensor order should be greater than zero.")
        if q == 0 or q > p:
            raise ValueError("Error calling ttvpy::ttv: contraction mode should be greater than zero or less than or equal to p.")
        A, (b,) = _torch_common(A, [b])
        return api.ttv(q, A, b)
    A = np.asarray(A)
    b = np.asarray(b)
    dt = _common_dtype(A, b)
    A = _as_array(A, dt)
    b = np.ascontiguousarray(b, dtype=dt)
    p = A.ndim
    if p == 0:
        raise ValueError("Error calling ttvpy::ttv: input tensor order should be greater than zero.")
    if q == 0 or q > p:
        raise ValueError("Error calling ttvpy::ttv: contraction mode should be greater than zero or less than or equal to p.")
    return np.ascontiguousarray(api.ttv(q, A, b))


def chain_plan(q: int, shape, order: str = "optimal"):
    """The sequence [(mode_to_contract, index_into_bs), ...] of the p-1 products, with modes renumbered after each
    contraction (wrapped_ttv.cpp:144-192).  Pure host logic."""
    p = len(shape)
    # vector j (0-based) belongs to original mode r: r = j+1 if j+1 < q else j+2
    modes = [(r, r - 1 if r < q else r - 2) for r in range(1, p + 1) if r != q]
    if order == "backward":
        seq = sorted(modes, key=lambda t: -t[0])
    elif order == "forward":
        seq = sorted(modes, key=lambda t: t[0])
    else:  # "optimal": the longest vector first, so that the tensor shrinks as fast as possible.  The reference sorts
        # ascending (stable for fewer than 16 vectors) and walks the list backwards (:170-188): ties -> larger mode first
        seq = sorted(modes, key=lambda t: int(shape[t[0] - 1]))[::-1]
    out = []
    alive = list(range(1, p + 1))           # original mode numbers still present, in order
    for r, j in seq:
        out.append((alive.index(r) + 1, j))
        alive.remove(r)
    return out


def ttvs(q: int, A, bs, order: str = "optimal"):
    """Multiplies A with p-1 vectors along every mode except q; returns the vector of length A.shape[q-1]."""
    if order not in ("optimal", "backward", "forward"):
        raise ValueError("Error calling ttvpy::ttvs: multiplication order should be either 'optimal', 'backward' or 'forward'.")
    on_device = api._is_torch(A)
    if not on_device:
        A = np.asarray(A)
    bs = [bj if api._is_torch(bj) else np.asarray(bj) for bj in bs]
    p = A.ndim
    if p == 0:
        raise ValueError("Error calling ttvpy::ttvs: input tensor order should be greater than zero.")
    if len(bs) != p - 1:
        raise ValueError("Error calling ttvpy::ttvs: number of input vectors is not equal to the tensor order - 1.")
    if q == 0 or q > p:
        raise ValueError("Error calling ttvpy::ttvs: contraction mode should be greater than zero or less than or equal to p.")
    if any(getattr(bj, "ndim", np.ndim(bj)) != 1 for bj in bs):
        raise ValueError("Error calling ttvpy::ttvs: some of the input vectors is not a vector.")
    shape = [int(s) for s in A.shape]
    want = [shape[r - 1] for r in range(1, p + 1) if r != q]
    if [int(bj.shape[0]) for bj in bs] != want:
        raise ValueError("Error calling ttvpy::ttvs: vector dimension is not compatible with the dimension of a tensor mode.")
    if on_device:
        A, bs = _torch_common(A, list(bs))
    else:
        dt = _common_dtype(A, *bs)                          # promote over A and every vector, never cast down to A's type
        A = _as_c_array(A, dt)
        bs = [np.ascontiguousarray(np.asarray(bj), dtype=dt) for bj in bs]
    if p == 1:
        return A

    if os.environ.get("TTV_B200_PY_CHAIN", "0") != "1":
        # the native chain (ttv_b200_ttvs): one call, intermediates in a stream-ordered pool in HBM, only c comes back
        last_order = list(range(p, 0, -1))                  # a C-contiguous array is a last-order tensor (wrapped_ttv.cpp:44-45)
        if on_device:
            return api.ttvs(q, A.contiguous(), shape, last_order, bs, order)
        return api.ttvs(q, A, shape, last_order, bs, order)
    return _ttvs_stepwise(q, A, bs, order, on_device, shape)


def _ttvs_stepwise(q, A, bs, order, on_device, shape):
    """the same chain as p-1 calls of the low-level interface from Python (TTV_B200_PY_CHAIN=1; also what CapturedTtvs
    records); the first product of a host tensor returns to the host before the rest continues on the device"""
    import torch
    steps = chain_plan(q, shape, order)
    if on_device:
        cur = A.contiguous()
        vecs = [bj.contiguous() for bj in bs]
    else:
        if not torch.cuda.is_available():
            raise api.TTVError(40, "Error in ttv_b200: CUDA failure (no CPU fallback exists). [no CUDA device]")
        # The first product reads all of A: it goes through the host-pointer path of the C-ABI, which streams A across
        # PCIe in chunks with the kernels overlapping the copies (a plain upload of pageable memory first measured
        # 11.5 GB/s on a 17 GB tensor).  Its result is n_q times smaller and continues on the device.
        mode, j = steps[0]
        first = api.ttv(mode, A, np.ascontiguousarray(np.asarray(bs[j]), dtype=A.dtype))
        steps = steps[1:]
        if not steps:
            return np.ascontiguousarray(first)
        cur = torch.from_numpy(np.ascontiguousarray(first)).cuda()
        vecs = [torch.from_numpy(np.ascontiguousarray(np.asarray(bj), dtype=A.dtype)).cuda() for bj in bs]
    for mode, j in steps:
        cur = api.ttv(mode, cur, vecs[j]).contiguous()      # output of a last-order tensor is last-order: no copy
    return cur if on_device else cur.cpu().numpy()


class CapturedTtvs:
    """The p-1 launches of ttvs(q, A, bs, order) on DEVICE tensors, captured once into a CUDA graph.

    On small tensors the chain is launch-bound: every product costs tens of microseconds of host time for a kernel
    that runs a few microseconds.  The asynchronous device-pointer path of the C-ABI only enqueues work on the caller's
    stream, so the whole chain can be recorded once and replayed with ONE launch.  The graph is bound to the storage of
    A and of the vectors: refill them in place, call replay(), read `result` (a device tensor owned by this object).

        plan = ttvpy.CapturedTtvs(q, A, bs)        # A: contiguous torch CUDA tensor, bs: p-1 CUDA vectors
        A.copy_(new_values); plan.replay(); y = plan.result
    """

    def __init__(self, q: int, A, bs, order: str = "optimal"):
        import torch
        if order not in ("optimal", "backward", "forward"):
            raise ValueError("Error calling ttvpy::ttvs: multiplication order should be either 'optimal', 'backward' or 'forward'.")
        if not (api._is_torch(A) and A.is_cuda and A.is_contiguous()):
            raise ValueError("CapturedTtvs needs a contiguous torch CUDA tensor (the graph is bound to its storage).")
        p = A.dim()
        if p < 2:
            raise ValueError("Error calling ttvpy::ttvs: input tensor order should be greater than one for a captured chain.")
        if q == 0 or q > p:
            raise ValueError("Error calling ttvpy::ttvs: contraction mode should be greater than zero or less than or equal to p.")
        if len(bs) != p - 1:
            raise ValueError("Error calling ttvpy::ttvs: number of input vectors is not equal to the tensor order - 1.")
        shape = [int(x) for x in A.shape]
        want = [shape[r - 1] for r in range(1, p + 1) if r != q]
        if [int(bj.shape[0]) for bj in bs] != want or any(bj.dim() != 1 for bj in bs):
            raise ValueError("Error calling ttvpy::ttvs: vector dimension is not compatible with the dimension of a tensor mode.")
        self.A, self.bs = A, [bj.contiguous() for bj in bs]
        self._steps = []                                   # (mode, input view, vector, flat output)
        cur, cur_shape = A, shape
        for mode, j in chain_plan(q, shape, order):
            out_shape = cur_shape[: mode - 1] + cur_shape[mode:]
            flat = torch.empty(int(np.prod(out_shape, dtype=object)), dtype=A.dtype, device=A.device)
            self._steps.append((mode, cur, self.bs[j], flat))
            cur, cur_shape = flat.view(*out_shape), out_shape   # a last-order tensor stays last-order: no copy
        self.result = cur
        self._stream = torch.cuda.Stream(A.device)
        self._stream.wait_stream(torch.cuda.current_stream(A.device))
        with torch.cuda.stream(self._stream):
            self._enqueue()                                # warm-up outside the capture: sizes workspaces
        self._stream.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, stream=self._stream):
            self._enqueue()

    def _enqueue(self):
        for mode, src, vec, flat in self._steps:
            api.ttv(mode, src, vec, out=flat, flags=api.FLAG_ASYNC, stream=self._stream)

    def replay(self):
        """recomputes `result` from the current contents of A and the vectors; ordered on the current stream"""
        self._graph.replay()
        return self.result
