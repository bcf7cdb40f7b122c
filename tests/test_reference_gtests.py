"""The reference's OWN GoogleTest sources (bassoy/ttv test/src/*.cpp), compiled unmodified against this repo's
include/tlib headers through the shim in tests/gtest_shim (see tests/ref_gtests.py).

  host binary: gtest_tlib_layout / shape / strides / workload  -- the restated L0 helpers                     (CPU)
  gpu binary:  gtest_tlib_ttv (all 19 policy combinations x {2,4,8}^p x all layouts x all q, double) and
               gtest_tlib_mtv (gemv_col / gemv_row / _parallel / _blas) -- every product runs on the B200     (GPU)

The binaries are built where /root/reference exists (the build container) and travel with the snapshot."""
from __future__ import annotations

import subprocess

import pytest

import ref_gtests


def _binary(name):
    path = ref_gtests.binary(name)
    if path is None and ref_gtests.reference_present():
        ref_gtests.build_all()
        path = ref_gtests.binary(name)
    if path is None:
        pytest.skip("tests/_refbin not built (needs /root/reference once)")
    return path


def test_reference_host_tests_pass_against_new_headers():
    r = subprocess.run([_binary("ref_gtests_host")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "[  PASSED  ] 18 tests." in r.stdout


@pytest.mark.gpu
def test_reference_ttv_and_mtv_tests_pass_on_the_gpu():
    r = subprocess.run([_binary("ref_gtests_gpu")], capture_output=True, text=True, timeout=3000)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "[  PASSED  ]" in r.stdout and "FAILED" not in r.stdout
