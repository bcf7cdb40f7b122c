"""The ttvpy drop-in (ttv_b200/ttvpy.py) against the reference's own Python tests (ttvpy/tests/test.py: einsum, exact
equality) and against fixtures produced by the reference's compiled module (tests/golden/ttvpy_golden.npz).  GPU."""
from __future__ import annotations

import os

import numpy as np
import pytest

import ttv_b200.ttvpy as tp

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ttvpy_golden.npz")


# ---- ttvpy/tests/test.py, restated ------------------------------------------------------------------------------------
def test_ttv_modes():
    A = np.arange(3 * 2 * 4, dtype=np.float64).reshape(3, 2, 4)
    for q, sub in ((1, "ijk,i->jk"), (2, "ijk,j->ik"), (3, "ijk,k->ij")):
        b = np.arange(A.shape[q - 1], dtype=np.float64)
        assert np.all(tp.ttv(q, A, b) == np.einsum(sub, A, b))


@pytest.mark.parametrize("q", [1, 2, 3, 4])
def test_ttvs_modes(q):
    for shape in ((3, 2, 4, 5), (5, 5, 5, 5)):
        A = np.arange(int(np.prod(shape)), dtype=np.float64).reshape(shape)
        B = [np.arange(shape[r], dtype=np.float64) for r in range(4) if r != q - 1]
        letters = "ijkl"
        D, rest, vecs = A, list(letters), list(B)
        for r in [x for x in range(4) if x != q - 1]:
            sub = "".join(rest) + "," + letters[r] + "->" + "".join(x for x in rest if x != letters[r])
            D = np.einsum(sub, D, vecs.pop(0))
            rest.remove(letters[r])
        for order in ("forward", "backward", "optimal"):
            C = tp.ttvs(q, A, B, order)
            assert C.shape == D.shape and np.all(C == D), (shape, q, order)


def test_golden_fixtures_of_reference_module():
    g = np.load(GOLDEN, allow_pickle=False)
    for m in range(int(g["count"])):
        A, b, q = g[f"ttv_A_{m}"], g[f"ttv_b_{m}"], int(g[f"ttv_q_{m}"])
        assert np.array_equal(tp.ttv(q, A, b), g[f"ttv_C_{m}"])
        bs = [g[f"ttvs_b_{m}_{j}"] for j in range(A.ndim - 1)]
        for order in ("forward", "backward", "optimal"):
            assert np.array_equal(tp.ttvs(q, A, bs, order), g[f"ttvs_C_{m}_{order}"]), (m, order)


def test_argument_errors_match_reference_texts():
    A = np.zeros((3, 2, 4))
    with pytest.raises(ValueError, match="contraction mode should be greater than zero"):
        tp.ttv(0, A, np.zeros(3))
    with pytest.raises(ValueError, match="contraction mode should be greater than zero"):
        tp.ttv(4, A, np.zeros(3))
    with pytest.raises(ValueError, match="multiplication order should be either"):
        tp.ttvs(1, A, [np.zeros(2), np.zeros(4)], "sideways")
    with pytest.raises(ValueError, match="number of input vectors"):
        tp.ttvs(1, A, [np.zeros(2)])
    with pytest.raises(ValueError, match="not compatible with the dimension"):
        tp.ttvs(1, A, [np.zeros(2), np.zeros(5)])
    with pytest.raises(ValueError, match="is not a vector"):
        tp.ttvs(1, A, [np.zeros((2, 1)), np.zeros(4)])


def test_other_dtypes_and_device_chain():
    import torch
    rng = np.random.default_rng(2)
    A = rng.integers(-3, 4, (4, 5, 3, 6)).astype(np.float32)
    bs = [rng.integers(-3, 4, n).astype(np.float32) for n in (4, 3, 6)]
    want = np.einsum("ijkl,i,k,l->j", A, *bs)
    assert np.array_equal(tp.ttvs(2, A, bs), want)
    got = tp.ttvs(2, torch.from_numpy(A).cuda(), [torch.from_numpy(b).cuda() for b in bs], "backward")
    assert got.is_cuda and np.array_equal(got.cpu().numpy(), want)
    Ai = A.astype(np.int32)
    assert np.array_equal(tp.ttv(3, Ai, bs[1].astype(np.int32)), np.einsum("ijkl,k->ijl", Ai, bs[1].astype(np.int32)))


def test_captured_chain_replays_on_new_data():
    """ttvpy.CapturedTtvs: the chain recorded into a CUDA graph gives what ttvs gives, also after A and the vectors were
    refilled in place (every q, every order, two element types)"""
    import torch
    from ttv_b200 import ttvpy
    rng = np.random.default_rng(5)
    for dtype in (np.float64, np.float32):
        shape = (6, 9, 4, 7)
        A = torch.from_numpy(rng.integers(-4, 5, shape).astype(dtype)).cuda()
        for q in range(1, 5):
            bs = [torch.from_numpy(rng.integers(-4, 5, shape[r]).astype(dtype)).cuda() for r in range(4) if r != q - 1]
            for order in ("optimal", "backward", "forward"):
                plan = ttvpy.CapturedTtvs(q, A, bs, order)
                for _ in range(2):
                    A.copy_(torch.from_numpy(rng.integers(-4, 5, shape).astype(dtype)))
                    for bj in bs:
                        bj.copy_(torch.from_numpy(rng.integers(-4, 5, bj.shape[0]).astype(dtype)))
                    got = plan.replay()
                    torch.cuda.synchronize()
                    letters = "abcd"
                    want = np.einsum(letters + "," + ",".join(letters[r] for r in range(4) if r != q - 1) + "->" + letters[q - 1],
                                     A.cpu().numpy(), *[bj.cpu().numpy() for bj in bs])
                    assert np.array_equal(got.cpu().numpy(), want), (dtype, q, order)


def _chain_reference(A, bs, q):
    """the chain by einsum, contracting one mode after the other (exact on small-integer data)"""
    p = A.ndim
    letters = "abcdefgh"[:p]
    return np.einsum(letters + "," + ",".join(letters[r] for r in range(p) if r != q - 1) + "->" + letters[q - 1], A, *bs)


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.complex64, np.complex128, np.int32, np.int64])
def test_native_chain_host_and_device(dtype, monkeypatch):
    """ttv_b200_ttvs (one native call, intermediates in the library's stream-ordered pool) against einsum and against the
    stepwise Python chain: orders 2..6, every q and order, extent-1 modes, host tensors (plain and chunked upload of A
    with the first result staying on the device) and device tensors; through a non-default layout as well"""
    import torch
    from ttv_b200 import api
    import ttv_b200
    rng = np.random.default_rng(11)
    shapes = [(7, 5), (1, 9), (6, 1, 4), (3, 2, 4, 5), (5, 1, 3, 1, 4), (4, 3, 2, 3, 2, 5), (64, 50, 83)]
    for shape in shapes:
        p = len(shape)
        A = rng.integers(-3, 4, shape).astype(dtype)
        if np.dtype(dtype).kind == "c":
            A = (A + 1j * rng.integers(-3, 4, shape)).astype(dtype)
        for q in range(1, p + 1):
            bs = [rng.integers(-2, 3, shape[r]).astype(dtype) for r in range(p) if r != q - 1]
            want = _chain_reference(A, bs, q)
            for order in ("optimal", "backward", "forward"):
                before = ttv_b200.launch_count()
                got = tp.ttvs(q, A, bs, order)
                assert got.shape == want.shape and np.array_equal(got, want), (shape, q, order, "host")
                assert ttv_b200.launch_count() - before >= p - 1
            monkeypatch.setenv("TTV_B200_PY_CHAIN", "1")
            assert np.array_equal(tp.ttvs(q, A, bs, "optimal"), want), (shape, q, "stepwise")
            monkeypatch.delenv("TTV_B200_PY_CHAIN")
            tA, tbs = torch.from_numpy(A).cuda(), [torch.from_numpy(b).cuda() for b in bs]
            for order in ("optimal", "forward"):
                got = tp.ttvs(q, tA, tbs, order)
                assert got.is_cuda and np.array_equal(got.cpu().numpy(), want), (shape, q, order, "device")
            # the same tensor stored first-order (Fortran order): the native entry takes any layout tuple
            flat = np.ascontiguousarray(A.reshape(-1, order="F"))
            got = api.ttvs(q, flat, list(shape), list(range(1, p + 1)), bs, "backward")
            assert np.array_equal(got, want), (shape, q, "first-order")
    # chunked upload: A of a few MB with 1 MiB chunks, the first product's result stays in HBM
    monkeypatch.setenv("TTV_B200_H2D_CHUNK_MB", "1")
    shape = (40, 30, 100, 9)
    A = rng.integers(-3, 4, shape).astype(dtype)
    for q in range(1, 5):
        bs = [rng.integers(-2, 3, shape[r]).astype(dtype) for r in range(4) if r != q - 1]
        want = _chain_reference(A, bs, q)
        for order in ("optimal", "backward", "forward"):
            assert np.array_equal(tp.ttvs(q, A, bs, order), want), (q, order, "chunked")


def test_native_chain_async_on_a_stream_and_errors():
    import torch
    from ttv_b200 import api
    import ttv_b200
    rng = np.random.default_rng(12)
    shape = (12, 7, 9, 5)
    A = rng.integers(-3, 4, shape).astype(np.float64)
    bs = [rng.integers(-2, 3, shape[r]).astype(np.float64) for r in (0, 1, 3)]
    want = _chain_reference(A, bs, 3)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        tA, tbs = torch.from_numpy(A).cuda(), [torch.from_numpy(b).cuda() for b in bs]
        out = torch.full((9,), -1.0, dtype=torch.float64, device="cuda")
        for _ in range(3):                                   # the pool hands the same blocks out again
            api.ttvs(3, tA, list(shape), [4, 3, 2, 1], tbs, "optimal", out=out, flags=api.FLAG_ASYNC, stream=stream)
    stream.synchronize()
    assert np.array_equal(out.cpu().numpy(), want)
    # argument errors keep the low-level interface's codes; host and device pointers must not be mixed
    with pytest.raises(ttv_b200.TTVError) as e:
        api.ttvs(5, A, list(shape), [4, 3, 2, 1], bs)
    assert e.value.status == 2
    with pytest.raises(ttv_b200.TTVError) as e:
        api.ttvs(3, A, list(shape), [4, 3, 3, 1], bs)
    assert e.value.status == 16
    with pytest.raises(ttv_b200.TTVError) as e:
        api.ttvs(3, A, list(shape), [4, 3, 2, 1], [bs[0], tbs[1], bs[2]])
    assert e.value.status == 41
    ttv_b200._lib.load().ttv_b200_release()                # waits, frees staging / workspaces, trims the pool
    assert np.array_equal(tp.ttvs(3, A, bs), want)         # and everything comes back on demand


def test_mixed_operand_types_are_promoted_not_truncated():
    """the reference binds array_t<double> and converts every operand (wrapped_ttv.cpp:205-206): an integer A with a fractional
    b, or a real A with a complex b, must not be cast down to A's type (round-1 advisor finding)"""
    rng = np.random.default_rng(11)
    A = rng.integers(-4, 5, (3, 2, 4))
    b = np.array([0.5, 0.25, -1.5])
    got = tp.ttv(1, A, b)
    assert got.dtype == np.float64 and np.allclose(got, np.einsum("ijk,i->jk", A, b))
    assert np.array_equal(tp.ttv(1, A, [1, 2, 3]), np.einsum("ijk,i->jk", A, [1, 2, 3]))          # integers stay exact
    Af = rng.uniform(-1, 1, (3, 2, 4)).astype(np.float32)
    bc = (rng.uniform(-1, 1, 2) + 1j * rng.uniform(-1, 1, 2)).astype(np.complex64)
    got = tp.ttv(2, Af, bc)
    assert got.dtype == np.complex64 and np.allclose(got, np.einsum("ijk,j->ik", Af, bc), atol=1e-6)
    bs = [np.array([0.5, 1.5]), np.array([1, 2, 3, 4])]                                          # float64 and int vectors, int tensor
    got = tp.ttvs(1, A, bs, "optimal")
    assert got.dtype == np.float64 and np.allclose(got, np.einsum("ijk,j,k->i", A, *bs))
    import torch
    tA = torch.from_numpy(A).cuda()                                                               # int64 device tensor, float vector
    got = tp.ttv(1, tA, torch.from_numpy(b).cuda())
    assert got.dtype == torch.float64 and np.allclose(got.cpu().numpy(), np.einsum("ijk,i->jk", A, b))
