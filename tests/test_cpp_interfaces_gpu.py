"""The two HIGH-LEVEL C++ interfaces BASELINE.json's north_star names, run on the GPU:

    (1) auto C = A(q) * b;                     reference include/tlib/ttv.h:122-127 (operator*), detail/tensor.h:90-95
    (2) auto C = ttv(q, A, b, ep, sp, fp);     reference include/tlib/ttv.h:99-114

* the reference's own example/interface{1,2,3}.cpp, compiled UNMODIFIED against include/ (tests/ref_gtests.py), must print
  the known answer {15,18,21,24,51,54,57,60} (interface1.cpp:43, interface2.cpp:43-44, interface3.cpp);
* tests/cpp/iface_check.cpp runs both interfaces on random order-3..5 tensors with non-trivial layouts for float / double /
  complex / int32 -- on plain host tensors, on a host tensor that keeps its copy in HBM (tensor::keep_on_device: the second
  product must not upload, mutable access must refresh the copy) and on device_tensor -- and every result is compared with
  the oracle and with the committed golden fixtures of the unmodified reference (tests/golden/ttv_golden.npz)."""
from __future__ import annotations

import os
import subprocess

import numpy as np
import pytest

import ref_gtests
from conftest import assert_close, random_case, real_case

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ttv_golden.npz")
CODES = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.complex64): 2, np.dtype(np.complex128): 3,
         np.dtype(np.int32): 4, np.dtype(np.int64): 5}


def _binary(name, needs_reference):
    path = ref_gtests.binary(name)
    if path is None:
        try:
            if needs_reference and not ref_gtests.reference_present():
                pytest.skip("tests/_refbin not built (needs /root/reference once)")
            ref_gtests.build_all() if needs_reference else ref_gtests.build_iface_check()
        except RuntimeError as exc:
            pytest.fail(str(exc))
        path = ref_gtests.binary(name)
    assert path is not None
    return path


@pytest.mark.parametrize("which", [1, 2, 3])
def test_reference_examples_run_unmodified_on_the_gpu(which):
    r = subprocess.run([_binary(f"ref_interface{which}", True)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    results = [line for line in r.stdout.splitlines() if line.startswith("C")]
    assert results, r.stdout
    for line in results:                                         # C1 (and C2 of interface 2)
        values = [float(x) for x in line.split("[")[1].split("]")[0].split()]
        assert values == [15, 18, 21, 24, 51, 54, 57, 60], line


def run_iface_many(tmp_path, cases):
    """cases: [(q, a, na, pia, b), ...] -> per case the six results of tests/cpp/iface_check.cpp (ONE process for all of them:
    a CUDA context per case would dominate the test time)"""
    case, out = tmp_path / "cases.bin", tmp_path / "out.bin"
    with open(case, "wb") as f:
        for q, a, na, pia, b in cases:
            np.array([CODES[a.dtype], len(na), q] + list(na) + list(pia), dtype=np.int64).tofile(f)
            np.ascontiguousarray(a).tofile(f)
            np.ascontiguousarray(b).tofile(f)
    r = subprocess.run([_binary("iface_check", False), str(case), str(out)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    raw = open(out, "rb").read()
    results, off = [], 0
    for q, a, na, pia, b in cases:
        n_out = a.size // na[q - 1]
        nbytes = 6 * n_out * a.dtype.itemsize
        results.append(np.frombuffer(raw, dtype=a.dtype, count=6 * n_out, offset=off).reshape(6, n_out))
        off += nbytes
    assert off == len(raw)
    return results


def run_iface(tmp_path, q, a, na, pia, b):
    return run_iface_many(tmp_path, [(q, a, na, pia, b)])[0]


CASES = [((7, 5, 6, 4), (2, 4, 1, 3)), ((6, 9, 5), (3, 1, 2)), ((4, 3, 5, 2, 6), (5, 2, 4, 1, 3)), ((8, 11, 7, 5), (4, 3, 2, 1)),
         ((33, 20, 17), (1, 2, 3))]


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128, np.int32])
def test_operator_and_tensor_level_interfaces_match_the_oracle(oracle, tmp_path, dtype):
    rng = np.random.default_rng(2026)
    cases = []
    for na, pia in CASES:
        for q in range(1, len(na) + 1):
            a, b = random_case(rng, na, q, dtype)
            cases.append((q, a, na, pia, b))
    for (q, a, na, pia, b), got in zip(cases, run_iface_many(tmp_path, cases)):
        want = oracle.ttv(q, a, na, pia, b)
        for i in range(5):
            assert np.array_equal(got[i], want), (na, pia, q, dtype, "variant", i)
        a2 = a.copy(); a2[0] += 1                                   # what iface_check did through A.begin()
        assert np.array_equal(got[5], oracle.ttv(q, a2, na, pia, b)), (na, pia, q, dtype, "after mutation")


def test_interfaces_on_real_valued_data_within_the_stated_tolerance(oracle, tmp_path):
    rng = np.random.default_rng(7)
    cases = []
    for dtype in (np.float32, np.complex128):
        for na, pia in CASES[:3]:
            for q in range(1, len(na) + 1):
                a, b = real_case(rng, na, q, dtype)
                cases.append((q, a, na, pia, b))
    for (q, a, na, pia, b), got in zip(cases, run_iface_many(tmp_path, cases)):
        ref, mag = oracle.naive(q, a, na, pia, b, want_abs=True)
        for i in range(5):
            assert_close(got[i], ref, mag, na[q - 1], a.dtype, what=f"{na} {pia} q={q} variant {i}")


def test_interfaces_reproduce_the_reference_fixtures(oracle, tmp_path):
    """tests/golden/ttv_golden.npz: outputs of the UNMODIFIED reference (tests/golden/make_golden.py)"""
    g = np.load(GOLDEN, allow_pickle=False)
    n_cases = int(g["count"])
    done = 0
    picked = []
    for i in range(0, n_cases, 6):                                   # every sixth case: all dtypes and orders come by
        na = [int(x) for x in g[f"na_{i}"]]; pia = [int(x) for x in g[f"pia_{i}"]]; q = int(g[f"q_{i}"])
        if len(na) >= 2:
            picked.append((i, q, g[f"a_{i}"], na, pia, g[f"b_{i}"]))
    for (i, q, a, na, pia, b), got in zip(picked, run_iface_many(tmp_path, [x[1:] for x in picked])):
        c = g[f"c_{i}"]
        for v in range(5):
            if bool(g[f"exact_{i}"]):
                assert np.array_equal(got[v], c), (i, na, pia, q, a.dtype, v)
            else:
                _, mag = oracle.naive(q, a, na, pia, b, want_abs=True)
                assert_close(got[v], c, mag, na[q - 1], a.dtype, what=f"golden case {i} variant {v}")
        done += 1
    assert done >= 30
