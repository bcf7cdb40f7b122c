"""Shared fixtures.  `-m "not gpu"` runs on a CPU-only box; `-m gpu` needs a B200 and goes through the C-ABI."""
from __future__ import annotations

import itertools
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu() -> bool:
    try:
        import ttv_b200
        return ttv_b200.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference compiled into oracle/_ref (None when it is not there, e.g. before build())."""
    from oracle.oracle import Reference
    return Reference() if Reference.available() else None


@pytest.fixture(scope="session")
def reference_blas():
    from oracle.oracle import Reference
    try:
        return Reference(blas=True) if Reference.available(blas=True) else None
    except OSError:
        return None


# ---- the reference's own test grid (test/src/gtest_tlib_ttv.cpp:192-425, test/include/gtest_aux.h:28-79) -------------
def reference_shapes(order: int, start: int = 2, steps: int = 3):
    """every shape in {start, 2*start, ...}^order, first extent fastest (gtest_aux.h:28-63)"""
    extents = [start * 2 ** j for j in range(steps)]
    return [tuple(reversed(s)) for s in itertools.product(extents, repeat=order)]


def all_layouts(order: int):
    """all order! layout tuples in lexicographic order (gtest_aux.h:65-79)"""
    return [tuple(pi) for pi in itertools.permutations(range(1, order + 1))]


def fold(na, pia, q):
    """(outer, nq, inner) of the canonical view; memory order of a packed (na, pia) tensor is [outer][nq][inner]."""
    k = list(pia).index(q)
    inner = int(np.prod([na[m - 1] for m in pia[:k]], dtype=object)) if k else 1
    outer = int(np.prod([na[m - 1] for m in pia[k + 1:]], dtype=object)) if k + 1 < len(pia) else 1
    return outer, int(na[q - 1]), inner


def reference_init(na, pia, q, dtype):
    """A as filled by ttv_init (gtest_tlib_ttv.cpp:28-71): the f-th mode-q fiber, fibers enumerated in layout order,
    holds f*nq+1 .. f*nq+nq.  In the canonical view that is A[o][k][i] = (o*inner + i)*nq + k + 1."""
    outer, nq, inner = fold(na, pia, q)
    f = np.arange(outer * inner, dtype=np.int64).reshape(outer, 1, inner)
    k = np.arange(nq, dtype=np.int64).reshape(1, nq, 1)
    return (f * nq + k + 1).astype(dtype).reshape(-1)


def reference_expected(na, q, n_out, dtype):
    """closed form of gtest_tlib_ttv.cpp:132 with b = 1: c[j] = (nq^2 (2(j+1) - 1) + nq) / 2 at memory offset j"""
    nq = int(na[q - 1])
    j = np.arange(1, n_out + 1, dtype=np.int64)
    return ((nq * nq * (2 * j - 1) + nq) // 2).astype(dtype)


def random_case(rng, na, q, dtype):
    """small-integer valued inputs: exact in every float type and free of signed overflow for ints"""
    n = int(np.prod(na, dtype=object))
    nq = int(na[q - 1])
    dt = np.dtype(dtype)
    if dt.kind == "c":
        a = (rng.integers(-8, 9, n) + 1j * rng.integers(-8, 9, n)).astype(dt)
        b = (rng.integers(-8, 9, nq) + 1j * rng.integers(-8, 9, nq)).astype(dt)
    else:
        a = rng.integers(-8, 9, n).astype(dt)
        b = rng.integers(-8, 9, nq).astype(dt)
    return a, b


def real_case(rng, na, q, dtype):
    """inputs in [-1, 1): rounding matters, compare with the n_q*eps tolerance"""
    n = int(np.prod(na, dtype=object))
    nq = int(na[q - 1])
    dt = np.dtype(dtype)
    if dt.kind == "c":
        a = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(dt)
        b = (rng.uniform(-1, 1, nq) + 1j * rng.uniform(-1, 1, nq)).astype(dt)
    elif dt.kind == "f":
        a = rng.uniform(-1, 1, n).astype(dt)
        b = rng.uniform(-1, 1, nq).astype(dt)
    else:
        return random_case(rng, na, q, dtype)
    return a, b


def eps_of(dtype) -> float:
    dt = np.dtype(dtype)
    return float(np.finfo(np.float32 if dt in (np.dtype(np.float32), np.dtype(np.complex64)) else np.float64).eps) / 2


def assert_close(c, c_ref, mag, nq, dtype, what=""):
    """The stated tolerance: |c - c_ref| <= 2 n_q eps sum_k |a_k||b_k| per element (complex: per component
    with the complex abs-sum); integers must be bit-exact."""
    dt = np.dtype(dtype)
    if dt.kind in "iu":
        assert np.array_equal(c, c_ref), f"integer result differs {what}"
        return
    tol = 2.0 * nq * eps_of(dt) * mag + 1e-300
    if dt.kind == "c":
        err = np.maximum(np.abs(c.real - c_ref.real), np.abs(c.imag - c_ref.imag))
    else:
        err = np.abs(c.astype(np.float64) - c_ref.astype(np.float64))
    bad = err > tol
    assert not bad.any(), f"{int(bad.sum())} of {c.size} elements outside 2*nq*eps*sum|a||b| {what}: max err {err.max()} tol {tol[bad].min()}"
