"""Round-1 advisor findings, pinned on the CPU (the GPU halves live in test_parity_gpu.py / test_ttvpy_gpu.py):

  * ttvpy promotes over ALL operands instead of casting b down to A's type (the reference binds array_t<double>,
    ttvpy/src/wrapped_ttv.cpp:205-206, and converts everything to float64)
  * api.ttv / ttv_lowlevel refuse operands whose element types, homes or density disagree (they travel as raw pointers)
  * the 'optimal' chain order breaks ties like the reference: ascending stable sort walked backwards (wrapped_ttv.cpp:170-188)
  * the library is rebuilt by CONTENT of every file under csrc/, atomically, under a lock
"""
import os

import numpy as np
import pytest

import ttv_b200
from ttv_b200 import api, build, ttvpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ttvpy_promotes_instead_of_casting_down():
    cd = ttvpy._common_dtype
    assert cd(np.zeros(2, np.int64), np.zeros(2, np.float64)) == np.float64          # integer A, fractional b: not truncated
    assert cd(np.zeros(2, np.float32), np.zeros(2, np.complex64)) == np.complex64      # real A, complex b: imaginary part kept
    assert cd(np.zeros(2, np.float64), np.zeros(2, np.complex64)) == np.complex128
    assert cd(np.zeros(2, np.float32), np.zeros(2, np.float32)) == np.float32          # agreeing operands keep their type
    assert cd(np.zeros(2, np.int32), np.zeros(2, np.int32)) == np.int32
    assert cd(np.zeros(2, np.int32), np.zeros(2, np.int64)) == np.int64
    assert cd(np.zeros(2, np.float16)) == np.float32 and cd(np.zeros(2, np.uint64)) == np.float64
    assert cd(np.zeros(2, bool), np.zeros(2, np.uint8)) == np.int32
    assert cd(np.zeros(2, np.float32), [0.5, 1.5]) == np.float64                        # python floats are float64, as in numpy


def test_operand_checks_fire_before_anything_reaches_the_device():
    a = np.zeros(24, np.float32); b = np.zeros(3, np.float64); c = np.zeros(8, np.float32)
    args = (2, 3, a, [4, 3, 2], [1, 4, 12], [1, 2, 3], b, [3], c, [4, 2], [1, 4], [1, 2])
    with pytest.raises(ttv_b200.TTVError) as err:
        ttv_b200.ttv_lowlevel(*args)
    assert err.value.status == 31 and "same element type" in str(err.value)
    b32 = np.zeros(6, np.float32)[::2]                                                  # strided b
    with pytest.raises(ttv_b200.TTVError) as err:
        ttv_b200.ttv_lowlevel(*(args[:6] + (b32,) + args[7:]))
    assert err.value.status == 32 and "dense" in str(err.value)
    A = np.zeros((4, 3, 2), np.float32)
    with pytest.raises(ttv_b200.TTVError) as err:                                       # out too small: would be written out of bounds
        api.ttv(2, A, np.zeros(3, np.float32), out=np.zeros(7, np.float32))
    assert err.value.status == 15
    with pytest.raises(ttv_b200.TTVError) as err:                                       # b must be a vector
        api.ttv(2, A, np.zeros((3, 1), np.float32))
    assert err.value.status == 13
    with pytest.raises(ttv_b200.TTVError) as err:                                       # dtype of b differs from A's
        api.ttv(2, A, np.zeros(3, np.float64))
    assert err.value.status == 31


def test_optimal_chain_order_breaks_ties_like_the_reference():
    # equal extents: the reference sorts ascending (stable below 16 vectors) and walks the list from the back -> larger mode first
    assert ttvpy.chain_plan(2, (5, 5, 5), "optimal") == [(3, 1), (1, 0)]
    assert ttvpy.chain_plan(1, (3, 4, 4, 4), "optimal") == [(4, 2), (3, 1), (2, 0)]
    assert ttvpy.chain_plan(3, (7, 7, 2, 7), "optimal") == [(4, 2), (2, 1), (1, 0)]
    for shape in [(5, 5, 5), (3, 4, 4, 4), (7, 7, 2, 7), (2, 9, 9, 2, 9), (6, 6, 6, 6, 6, 6)]:
        for q in range(1, len(shape) + 1):
            for order in ("optimal", "backward", "forward"):
                assert api.chain_plan(q, shape, order) == ttvpy.chain_plan(q, shape, order), (shape, q, order)


def test_library_staleness_is_decided_by_content_of_every_source():
    srcs = build.sources()
    names = {os.path.basename(f) for f in srcs}
    on_disk = {f for f in os.listdir(os.path.join(ROOT, "ttv_b200", "csrc")) if os.path.isfile(os.path.join(ROOT, "ttv_b200", "csrc", f))}
    assert on_disk <= names, on_disk - names                     # a new kernel header can never be forgotten (scatter_kernel.cuh was)
    assert {"scatter_kernel.cuh", "colt_kernel.cuh", "api.cu", "plan.cpp", "ttv_b200.h", "build.py"} <= names
    assert not build.stale()                                       # the library in the tree matches the sources ...
    stamp = open(build.STAMP).read().strip()
    assert stamp == build.source_hash() and len(stamp) == 64
    os.utime(os.path.join(ROOT, "ttv_b200", "csrc", "plan.cpp"))   # ... and a fresh mtime (a snapshot copied elsewhere) changes nothing
    assert not build.stale()


def test_roofline_traffic_file_names_the_kernel_the_chooser_launches():
    """bench.py reports `roofline.traffic` only while profiles/roofline_traffic.json was captured for the kernel this tree
    launches on the bench workload (the number must not go stale silently when the chooser changes)"""
    import json
    rec = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    for q in (2, 3, 4):
        pl = ttv_b200.plan(q, [256] * 4, [1, 2, 3, 4], dtype="f32")
        assert pl["kernel"] == 2 and rec["kernel"] == f"ttv_col_kernel<float,{pl['vec']},{pl['nu']},{pl['ku']}>"
    assert rec["algorithmic_bytes_per_launch"] == 4 * (256 ** 4 + 256 + 256 ** 3)
    assert 0.99 < rec["col_kernel_dram_bytes_per_launch"] / rec["algorithmic_bytes_per_launch"] < 1.02
