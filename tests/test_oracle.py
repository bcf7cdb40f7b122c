"""Pins the CPU oracle (oracle/ttv_oracle.c) before anything trusts it:
  1. the reference's own known answers (closed forms of its gtest suite, its examples, ttvpy's tests and README),
  2. the committed golden fixtures produced by the unmodified reference (tests/golden/make_golden.py),
  3. the unmodified reference itself, live, when oracle/_ref has been built (all 17 policy combinations, both builds).
CPU only."""
from __future__ import annotations

import os

import numpy as np
import pytest

from conftest import (all_layouts, fold, random_case, real_case, reference_expected, reference_init,
                      reference_shapes, assert_close)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ttv_golden.npz")


# ---- 1. known answers of the reference ---------------------------------------------------------------------------
@pytest.mark.parametrize("order", [2, 3, 4])
@pytest.mark.parametrize("slicing", ["slice", "subtensor"])
def test_reference_grid_closed_form(oracle, order, slicing):
    """gtest_tlib_ttv.cpp:192-425: double, every shape in {2,4,8}^p, all p! layouts, every q; expected :132."""
    for na in reference_shapes(order):
        for pia in all_layouts(order):
            for q in range(1, order + 1):
                a = reference_init(na, pia, q, np.float64)
                b = np.ones(na[q - 1], np.float64)
                c = oracle.ttv(q, a, na, pia, b, slicing)
                assert np.array_equal(c, reference_expected(na, q, c.size, np.float64)), (na, pia, q)


def test_mtv_closed_form(oracle):
    """gtest_tlib_mtv.cpp:100-125: matrices {2..1024}^2, a(i,j) = i*n + j + 1, b = 1; row sums :70-76."""
    for m in [2 ** e for e in range(1, 11)]:
        for n in [2 ** e for e in range(1, 11, 3)]:
            i = np.arange(m).reshape(m, 1); j = np.arange(n).reshape(1, n)
            vals = (i * n + j + 1).astype(np.float64)
            ii = np.arange(1, m + 1, dtype=np.int64)
            fn = lambda t: (t * n * (t * n + 1)) // 2
            expect = (fn(ii) - fn(ii - 1)).astype(np.float64)
            col_major = np.ascontiguousarray(vals.T).reshape(-1)      # a[i + j*m]
            row_major = vals.reshape(-1)                              # a[j + i*n]
            b = np.ones(n)
            assert np.array_equal(oracle.gemv("col", col_major, b, m, n, m), expect)
            assert np.array_equal(oracle.gemv("row", row_major, b, m, n, n), expect)


def test_example_known_answer(oracle):
    """example/interface{1,2,3}.cpp: A = iota(1..24), shape (4,3,2), first-order, b = 1, q = 2."""
    a = np.arange(1, 25, dtype=np.float32)
    for slicing in ("slice", "subtensor"):
        c = oracle.ttv(2, a, [4, 3, 2], [1, 2, 3], np.ones(3, np.float32), slicing)
        assert c.tolist() == [15, 18, 21, 24, 51, 54, 57, 60]


def test_ttvpy_known_answers(oracle):
    """ttvpy/tests/test.py:6-27 (einsum) and ttvpy/README.md:61-66; numpy C order = last-order layout."""
    A = np.arange(24, dtype=np.float64).reshape(3, 2, 4)
    pia = [3, 2, 1]
    for q, sub in ((1, "ijk,i->jk"), (2, "ijk,j->ik"), (3, "ijk,k->ij")):
        b = np.arange(A.shape[q - 1], dtype=np.float64)
        c = oracle.ttv(q, A.reshape(-1), A.shape, pia, b)
        assert np.array_equal(c, np.einsum(sub, A, b).reshape(-1))
    c = oracle.ttv(1, A.reshape(-1), A.shape, pia, np.arange(3, dtype=np.float64))
    assert c.reshape(2, 4).tolist() == [[40, 43, 46, 49], [52, 55, 58, 61]]


# ---- 2. committed golden fixtures (outputs of the unmodified reference) --------------------------------------------
def test_golden_fixtures(oracle):
    assert os.path.exists(GOLDEN), "tests/golden/ttv_golden.npz is missing (run tests/golden/make_golden.py)"
    g = np.load(GOLDEN, allow_pickle=False)
    n = int(g["count"])
    assert n > 100
    for i in range(n):
        na = g[f"na_{i}"].tolist(); pia = g[f"pia_{i}"].tolist(); q = int(g[f"q_{i}"])
        a, b, c_ref = g[f"a_{i}"], g[f"b_{i}"], g[f"c_{i}"]
        for slicing in ("slice", "subtensor"):
            c = oracle.ttv(q, a, na, pia, b, slicing)
            if bool(g[f"exact_{i}"]):
                assert np.array_equal(c, c_ref), (i, na, pia, q, a.dtype)
            else:
                _, mag = oracle.naive(q, a, na, pia, b, want_abs=True)
                assert_close(c, c_ref, mag, na[q - 1], a.dtype, what=f"golden case {i}")


# ---- 3. the unmodified reference, live -------------------------------------------------------------------------------
def test_live_reference_all_policies(oracle, reference, reference_blas):
    if reference is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from oracle.oracle import REF_COMBOS
    rng = np.random.default_rng(7)
    libs = [reference] + ([reference_blas] if reference_blas is not None else [])
    for order in (2, 3, 4):
        for na in [(2, 3, 4, 5)[:order], (4, 2, 3, 2)[:order], (3, 5, 2, 4)[:order]]:
            for pia in all_layouts(order):
                for q in range(1, order + 1):
                    for dtype in (np.float32, np.float64, np.int32, np.int64, np.complex64, np.complex128):
                        a, b = random_case(rng, na, q, dtype)
                        c = oracle.ttv(q, a, na, pia, b)
                        assert np.array_equal(c, oracle.naive(q, a, na, pia, b))
                        for lib in libs:
                            for combo in REF_COMBOS:
                                assert np.array_equal(c, lib.ttv(q, a, na, pia, b, combo=combo)), (na, pia, q, dtype, combo)


def test_live_reference_rounding(oracle, reference):
    """Non-integer data: the oracle's sequential path is bit-identical to the reference's sequential path compiled
    without FMA contraction differences only up to the stated tolerance; the naive wide checker bounds both."""
    if reference is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(11)
    for na, pia in [((7, 33, 5), (1, 2, 3)), ((7, 33, 5), (3, 1, 2)), ((6, 5, 129), (2, 3, 1)), ((17, 300), (1, 2)), ((17, 300), (2, 1))]:
        for q in range(1, len(na) + 1):
            for dtype in (np.float32, np.float64, np.complex64, np.complex128):
                a, b = real_case(rng, na, q, dtype)
                wide, mag = oracle.naive(q, a, na, pia, b, want_abs=True)
                c = oracle.ttv(q, a, na, pia, b)
                r = reference.ttv(q, a, na, pia, b, combo=("par_loop", "subtensor", "all"))
                assert_close(c, wide, mag, na[q - 1], dtype, "oracle vs wide")
                assert_close(r, wide, mag, na[q - 1], dtype, "reference vs wide")


def test_tensor_interface_of_reference(oracle, reference):
    """interfaces 1 and 2 (ttv.h:99-127) produce the output shape / layout / strides the helpers predict"""
    if reference is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    a = np.arange(1, 25, dtype=np.float32)
    for use_op in (True, False):
        c, nc, pic, wc = reference.tensor_iface(2, a, [4, 3, 2], [1, 2, 3], np.ones(3, np.float32), use_op)
        assert c.tolist() == [15, 18, 21, 24, 51, 54, 57, 60]
        assert nc == oracle.output_shape([4, 3, 2], 2) and pic == oracle.output_layout([1, 2, 3], 2)
        assert wc == oracle.strides(nc, pic)


def test_fold_matches_memory_order(oracle):
    """the canonical view used by every test helper: flat memory of a packed (na, pia) tensor is [outer][nq][inner]"""
    rng = np.random.default_rng(3)
    for na, pia in [((3, 4, 5), (2, 3, 1)), ((2, 3, 4, 5), (4, 1, 3, 2)), ((5, 2), (2, 1))]:
        for q in range(1, len(na) + 1):
            outer, nq, inner = fold(na, pia, q)
            a, b = random_case(rng, na, q, np.int64)
            view = a.reshape(outer, nq, inner)
            expect = np.einsum("okj,k->oj", view, b).reshape(-1)
            assert np.array_equal(oracle.ttv(q, a, na, pia, b), expect)


# ---- 5. the chain of p-1 products (ttvpy::ttvs) ---------------------------------------------------------------------------
def test_chain_schedule_with_the_oracle_reproduces_the_reference_module(oracle):
    """The schedule ttv_b200_ttvs follows (ttv_b200_chain_plan, pure host code) executed step by step with the ORACLE's
    ttv on last-order tensors must reproduce what the reference's own compiled module returned for ttvs(q, A, bs, order)
    -- tests/golden/ttvpy_golden.npz -- for all three orders (wrapped_ttv.cpp:135-192): pins the schedule on the CPU."""
    import ttv_b200
    from ttv_b200 import api
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ttvpy_golden.npz")
    g = np.load(path, allow_pickle=False)
    checked = 0
    for m in range(int(g["count"])):
        A, q = g[f"ttv_A_{m}"], int(g[f"ttv_q_{m}"])
        p = A.ndim
        if p < 2:
            continue
        bs = [g[f"ttvs_b_{m}_{j}"] for j in range(p - 1)]
        for order in ("forward", "backward", "optimal"):
            cur, shape = np.ascontiguousarray(A).reshape(-1), list(A.shape)
            for mode, j in api.chain_plan(q, list(A.shape), order):
                pia = list(range(len(shape), 0, -1))                    # C-contiguous = last-order
                cur = oracle.ttv(mode, cur, shape, pia, np.ascontiguousarray(bs[j], dtype=cur.dtype))
                del shape[mode - 1]
            want = g[f"ttvs_C_{m}_{order}"]
            assert cur.shape == want.reshape(-1).shape, (m, order)
            # the fixtures hold small integers in float64: every summation order is exact
            assert np.array_equal(cur, want.reshape(-1)), (m, q, order)
            checked += 1
    assert checked >= 30
