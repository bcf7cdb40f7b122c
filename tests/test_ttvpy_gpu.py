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
