"""CPU-only tests of the boundary: the C-ABI library loads and exports every symbol include/ttv_b200.h declares, the
restated L0 helpers agree with the oracle (and with the live reference when built), the argument checks fire in the
reference's order with its messages, and the layout folder / kernel chooser behave.  No compute calls (no GPU here)."""
from __future__ import annotations

import ctypes as C
import itertools
import os
import re

import numpy as np
import pytest

import ttv_b200
from ttv_b200 import _lib
from conftest import all_layouts, fold

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ttv_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(ttv_b200_[a-z0-9_]+)\s*\(", body))
    assert len(declared) >= 25
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/ttv_b200.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.ttv_b200_version() == 110


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.Opts) == 40
    assert C.sizeof(_lib.Plan) == 3 * 8 + 2 * 4 + 10 * 4 + 5 * 8


# ---- L0 helpers ------------------------------------------------------------------------------------------------------
def test_layout_helpers_match_reference_literals():
    """literals of test/src/gtest_tlib_layout.cpp and gtest_tlib_strides.cpp / gtest_tlib_shape.cpp"""
    g = ttv_b200.generate_k_order_layout
    assert g(4, 1) == [1, 2, 3, 4] and g(4, 2) == [2, 1, 3, 4] and g(4, 3) == [3, 2, 1, 4]
    assert g(4, 4) == [4, 3, 2, 1] and g(4, 0) == [4, 3, 2, 1] and g(4, 9) == [4, 3, 2, 1] and g(1, 1) == [1]
    for bad in ([], [0], [0, 1], [1, 0], [2], [1, 1], [1, 3], [1, 3, 5], [2, 2, 1]):
        assert not ttv_b200.is_valid_layout(bad), bad
    for good in ([1], [1, 2], [2, 1], [3, 1, 2], [4, 3, 2, 1]):
        assert ttv_b200.is_valid_layout(good), good
    s = ttv_b200.generate_strides
    assert s([4, 4, 2], [1, 2, 3]) == [1, 4, 16]
    assert s([4, 2, 2], [2, 1, 3]) == [2, 1, 8]
    assert s([4, 4, 2], [3, 2, 1]) == [8, 2, 1]
    assert s([1, 1], [1, 2]) == [1, 1] and s([4, 1], [1, 2]) == [1, 1] and s([1, 4], [2, 1]) == [1, 1]
    assert s([4], [1]) == [1] and s([1], [1]) == [1]
    assert s([3, 4], [1, 2]) == [1, 3] and s([3, 4], [2, 1]) == [4, 1]
    assert ttv_b200.generate_output_shape([4, 3, 2], 1) == [3, 2]
    assert ttv_b200.generate_output_shape([4, 3, 2], 2) == [4, 2]
    assert ttv_b200.generate_output_shape([4, 3, 2], 3) == [4, 3]
    assert ttv_b200.generate_output_layout([3, 1, 2], 1) == [2, 1]
    assert ttv_b200.generate_output_layout([3, 1, 2], 3) == [1, 2]
    assert not ttv_b200.is_valid_shape([]) and not ttv_b200.is_valid_shape([3, 0]) and ttv_b200.is_valid_shape([1])


def test_helpers_agree_with_oracle_and_reference(oracle, reference):
    shapes = [s for p in (1, 2, 3, 4) for s in itertools.product([1, 2, 3], repeat=p)]
    for n in shapes:
        p = len(n)
        for pi in all_layouts(p):
            w = ttv_b200.generate_strides(n, pi)
            assert w == oracle.strides(n, pi)
            assert ttv_b200.is_valid_strides(pi, w) == oracle.is_valid_strides(pi, w)
            if reference is not None:
                rw = np.zeros(p, np.uint64)
                nn, pp = np.asarray(n, np.uint64), np.asarray(pi, np.uint64)
                reference.lib.ttv_ref_compute_strides(nn.ctypes.data_as(_lib.u64p), pp.ctypes.data_as(_lib.u64p), p,
                                                      rw.ctypes.data_as(_lib.u64p))
                assert w == [int(x) for x in rw], (n, pi)
            for q in range(1, p + 1):
                if p > 1:
                    assert ttv_b200.generate_output_shape(n, q) == oracle.output_shape(n, q)
                    assert ttv_b200.generate_output_layout(pi, q) == oracle.output_layout(pi, q)
    for p in range(1, 8):
        for k in range(0, p + 2):
            assert ttv_b200.generate_k_order_layout(p, k) == oracle.k_order_layout(p, k)
    # strides that decrease along the layout are invalid (strides.h:76-101)
    assert not ttv_b200.is_valid_strides([1, 2, 3], [1, 8, 4])
    assert ttv_b200.is_valid_strides([1, 2, 3], [1, 4, 4])


# ---- argument checks ---------------------------------------------------------------------------------------------------
def _plan_status(q, p, a, na, wa, pia, b, nb, c, nc, wc, pic):
    lib = _lib.load()
    keep = [ttv_b200.api._tuple(v) for v in (na, wa, pia, nb, nc, wc, pic)]
    ptrs = [k[1] for k in keep]
    st = lib.ttv_b200_plan(0, q, p, a, ptrs[0], ptrs[1], ptrs[2], b, ptrs[3], c, ptrs[4], ptrs[5], ptrs[6], None, None)
    return st, lib.ttv_b200_last_error().decode() if st else ""


def test_argument_checks_follow_the_reference(oracle, reference):
    """ttv.h:64-89: sixteen checks, evaluated in order; same status and same text as the oracle (and as the what()
    of the exception the live reference throws)."""
    one = C.c_void_p(64)
    ok = dict(q=2, p=3, a=one, na=[4, 3, 2], wa=[1, 4, 12], pia=[1, 2, 3], b=one, nb=[3], c=one, nc=[4, 2], wc=[1, 4], pic=[1, 2])
    assert _plan_status(**ok)[0] == 0
    cases = [
        (1, dict(p=0)), (2, dict(q=0)), (2, dict(q=4)), (3, dict(a=None)), (4, dict(b=None)), (5, dict(c=None)),
        (6, dict(na=None)), (7, dict(nb=None)), (8, dict(nc=None)), (9, dict(wa=None)), (10, dict(wc=None)),
        (11, dict(pia=None)), (12, dict(pic=None)), (13, dict(nb=[4])), (14, dict(na=[4, 3, 0])), (15, dict(nc=[0, 2])),
        (16, dict(pia=[1, 2, 2])), (16, dict(pia=[0, 1, 2])), (17, dict(pic=[1, 1])), (17, dict(pic=[3, 1])),
        (18, dict(wa=[12, 4, 1])), (19, dict(wc=[4, 1])),
        (20, dict(pic=[2, 1], wc=[2, 1])),                                 # beginning of the layout tuples differs (case 8 only)
        (2, dict(p=0 + 1, q=2)),                                 # q > p
        (15, dict(p=1, q=1, na=[3], wa=[1], pia=[1], nc=[1], wc=[1], pic=[1])),   # order 1 always throws here
        # several things wrong at once: the earliest check wins
        (3, dict(a=None, na=None, pia=[9, 9, 9])), (13, dict(nb=[9], na=[4, 3, 0])), (14, dict(na=[0, 3, 2], pia=[1, 1, 1])),
    ]
    dummy = np.zeros(64)
    for want, change in cases:
        args = dict(ok); args.update(change)
        st, msg = _plan_status(**args)
        assert st == want, (want, change, st, msg)
        assert msg.startswith(oracle.strerror(want)), (msg, oracle.strerror(want))
        # the oracle, given the same arguments, reports the same status
        def arr(v):
            return None if v is None else np.asarray(v, np.uint64)
        def buf(v):
            return None if v is None else dummy
        ost = oracle.run_raw(1, 1, args["q"], args["p"], buf(args["a"]), arr(args["na"]), arr(args["wa"]), arr(args["pia"]),
                             buf(args["b"]), arr(args["nb"]), buf(args["c"]), arr(args["nc"]), arr(args["wc"]), arr(args["pic"]))
        assert ost == want, (want, change, ost)
        if reference is not None:
            rst = reference.run_raw(1, ("seq", "slice", "none"), args["q"], args["p"], buf(args["a"]), arr(args["na"]),
                                    arr(args["wa"]), arr(args["pia"]), buf(args["b"]), arr(args["nb"]), buf(args["c"]),
                                    arr(args["nc"]), arr(args["wc"]), arr(args["pic"]))
            assert rst == -1 and reference.last_error() == oracle.strerror(want), (want, change, reference.last_error())


def test_layout_end_mismatch():
    one = C.c_void_p(64)
    st, msg = _plan_status(2, 4, one, [2, 3, 4, 5], [1, 2, 6, 24], [1, 2, 3, 4], one, [3], one, [2, 4, 5], [1, 10, 2], [1, 3, 2])
    assert st == 21 and "end of layout tuples" in msg


def test_non_packed_strides_take_the_general_stride_kernel_in_case_8_only():
    one = C.c_void_p(64)
    # padded leading dimension of A (ld 8 for 4 rows), packed C: honoured like the reference's slice variants do
    st, msg = _plan_status(2, 3, one, [4, 3, 2], [1, 8, 24], [1, 2, 3], one, [3], one, [4, 2], [1, 4], [1, 2])
    assert st == 0, msg
    pl = ttv_b200.plan(2, [4, 3, 2], [1, 2, 3], dtype="f64", wa=[1, 8, 24])
    assert pl["kernel"] == 6 and (pl["outer"], pl["nq"], pl["inner"]) == (2, 3, 4)
    assert ttv_b200.plan(2, [4, 3, 2], [1, 2, 3], dtype="f64", wc=[1, 6])["kernel"] == 6
    assert ttv_b200.plan(2, [4, 3, 2], [1, 2, 3], dtype="f64")["kernel"] == 2
    # C's shape must be A's without mode q when the strides have to be matched up
    st, msg = _plan_status(2, 3, one, [4, 3, 2], [1, 8, 24], [1, 2, 3], one, [3], one, [4, 3], [1, 4], [1, 2])
    assert st == 30, msg
    # cases 1-7 ignore wa / wc exactly like the reference's mtv (matrix_times_vector.h:314-336)
    st, _ = _plan_status(1, 3, one, [4, 3, 2], [1, 8, 24], [1, 2, 3], one, [4], one, [3, 2], [1, 3], [1, 2])
    assert st == 0
    # extent-1 modes may carry any stride
    st, _ = _plan_status(2, 4, one, [3, 2, 1, 4], [1, 3, 4, 6], [1, 2, 3, 4], one, [2], one, [3, 1, 4], [1, 2, 3], [1, 2, 3])
    assert st == 0


# ---- folder + chooser ----------------------------------------------------------------------------------------------------
def test_fold_and_case_classification(oracle):
    for p in (2, 3, 4, 5):
        for na in [(2, 3, 4, 5, 6)[:p], (5, 1, 3, 1, 2)[:p]]:
            for pia in all_layouts(p):
                for q in range(1, p + 1):
                    pl = ttv_b200.plan(q, na, pia, dtype="f64")
                    assert (pl["outer"], pl["nq"], pl["inner"]) == fold(na, pia, q)
                    assert pl["ref_case"] == oracle.case(p, q, pia)
                    assert pl["k"] == list(pia).index(q) + 1
                    n = int(np.prod(na))
                    assert pl["algo_bytes"] == 8 * (n + na[q - 1] + n // na[q - 1])
                    assert pl["algo_flops"] == 2 * n
                    assert pl["kernel"] in (((9,) if pl["nq"] == 2 else (1,)) if pl["inner"] == 1 else (2, 10))


def test_chooser_named_configs():
    # cfg1: 512^3 fp32 q=2 -> column GEMV, 16-byte vectors, a full row per CTA
    pl = ttv_b200.plan(2, [512, 512, 512], [1, 2, 3], dtype="f32")
    assert pl["kernel"] == 2 and pl["vec"] == 4 and pl["tx"] == 128 and pl["algo_bytes"] == 537921536
    # cfg5 q=1: DOT, one warp per 16 KiB fiber; q=3: column GEMV over 4 Mi outputs
    pl = ttv_b200.plan(1, [2048] * 3, [1, 2, 3], dtype="f64")
    assert pl["kernel"] == 1 and pl["vec"] == 2 and pl["ty"] == 32 and pl["ksplit"] == 1
    pl = ttv_b200.plan(3, [2048] * 3, [1, 2, 3], dtype="f64")
    assert pl["kernel"] == 2 and pl["vec"] == 2 and pl["tx"] == 256 and pl["ksplit"] == 1
    # odd inner extent: scalar loads along inner; odd n_q: scalar loads along n_q
    assert ttv_b200.plan(2, [23, 23, 23], [1, 2, 3], dtype="f32")["vec"] == 1
    assert ttv_b200.plan(1, [23, 23, 23], [1, 2, 3], dtype="f32")["vec"] == 1
    # odd WIDE rows: phase lanes along n_q keep 16-byte loads (COLX)
    pl = ttv_b200.plan(2, [1625] * 3, [1, 2, 3], dtype="f32")
    assert pl["kernel"] == 4 and pl["vec"] == 4 and pl["ty"] == 4 and pl["tx"] * 4 <= 256
    pl = ttv_b200.plan(7, [21] * 7, [1, 2, 3, 4, 5, 6, 7], dtype="f64")        # short n_q: one warp per 31 vectors
    assert pl["kernel"] == 4 and pl["vec"] == 2 and (pl["tx"], pl["ty"]) == (32, 1)
    pl = ttv_b200.plan(3, [215] * 4, [1, 2, 3, 4], dtype="f64")
    assert pl["kernel"] == 4 and pl["vec"] == 2 and pl["ty"] == 2
    assert ttv_b200.plan(2, [1625] * 3, [1, 2, 3], dtype="c128")["kernel"] == 2
    with pytest.raises(ttv_b200.TTVError):
        ttv_b200.plan(2, [64, 64, 64], [1, 2, 3], dtype="c128", kernel="colx")
    # many short aligned fibers: one flat stream (DOTF); long ones keep a lane group per fiber (DOT)
    assert ttv_b200.plan(1, [40] * 6, [1, 2, 3, 4, 5, 6], dtype="f32")["kernel"] == 5
    assert ttv_b200.plan(1, [256] * 4, [1, 2, 3, 4], dtype="f32")["kernel"] == 1
    # tiny extents of the asymmetric family ("Tiny extents"): fibers of two elements -> DOTP, a CTA per tile
    pl = ttv_b200.plan(1, [2, 2, 4, 2, 1 << 15, 2, 3, 2, 2, 128], list(range(1, 11)), dtype="i32")
    assert (pl["kernel"], pl["vec"], pl["ku"], pl["ctas"]) == (9, 4, 8, 1610612736 // 2 // 2048)
    assert ttv_b200.plan(1, [2, 1 << 20], [1, 2], dtype="f64")["kernel"] == 9 and ttv_b200.plan(1, [2, 1 << 20], [1, 2], dtype="c128")["kernel"] != 9
    # rows that are not whole vectors -> COLF: long slabs cut over warps, super-rows of two rows (rows of 6) / four rows (rows of 3)
    pl = ttv_b200.plan(3, [2, 3, 1 << 20, 2, 4, 16], [1, 2, 3, 4, 5, 6], dtype="f32")
    assert (pl["kernel"], pl["tx"], pl["ty"], pl["to"], pl["nu"]) == (10, 3, 10, 2, 1) and 50 <= pl["ksplit"] <= 3552 and pl["ctas"] == 148 * 6
    pl = ttv_b200.plan(2, [3, 1 << 20, 64], [1, 2, 3], dtype="f32")
    assert (pl["kernel"], pl["tx"], pl["ty"], pl["to"]) == (10, 3, 10, 4) and pl["workspace_bytes"] == pl["ksplit"] * 64 * 3 * 4
    assert ttv_b200.plan(2, [3, 1 << 26], [1, 2], dtype="f32")["ksplit"] <= 148 * 24                    # one slab: at most 24 partitions per SM
    # short slabs side by side in a warp (8 388 608 slabs of 128 x 2: eight lanes per slab, four slabs per warp, 32 CTAs per SM)
    pl = ttv_b200.plan(2, [2, 128, 2, 2, 1 << 21], [1, 2, 3, 4, 5], dtype="f32")
    assert (pl["kernel"], pl["tx"], pl["ty"], pl["to"], pl["nu"], pl["ksplit"], pl["ctas"]) == (10, 1, 8, 2, 4, 1, 148 * 32)
    pl = ttv_b200.plan(2, [5, 64, 1 << 20], [1, 2, 3], dtype="f32")
    assert (pl["kernel"], pl["tx"], pl["ty"], pl["to"], pl["nu"], pl["ksplit"], pl["ctas"]) == (10, 5, 2, 4, 3, 1, 148 * 6)
    # tiny slabs of two-element rows: consecutive lanes on consecutive vectors, a CTA per eight items
    pl = ttv_b200.plan(2, [2, 2, 4, 2, 1 << 15, 2, 3, 2, 2, 128], list(range(1, 11)), dtype="f32")
    assert (pl["kernel"], pl["ty"], pl["nu"], pl["smem_bytes"], pl["ctas"]) == (10, 1, 32, 0, 805306368 // 256 // 8)
    assert ttv_b200.plan(2, [2, 16, 1 << 20], [1, 2, 3], dtype="i32")["ty"] == 8 and ttv_b200.plan(2, [2, 16, 1 << 20], [1, 2, 3], dtype="f64")["kernel"] == 2
    # 8-byte elements: COLF for rows of up to 8 vectors only; rows of whole vectors never
    assert ttv_b200.plan(2, [3, 1 << 19, 64], [1, 2, 3], dtype="f64")["kernel"] == 10 and ttv_b200.plan(2, [21, 1 << 16, 128], [1, 2, 3], dtype="f64")["kernel"] == 2
    assert ttv_b200.plan(2, [4, 1 << 20, 64], [1, 2, 3], dtype="f32")["kernel"] == 2
    # several slabs of odd 4-byte rows off the 16-byte grid: the one place STREAMK is taken by rule
    assert ttv_b200.plan(2, [3, (1 << 20) + 1, 64], [1, 2, 3], dtype="f32")["kernel"] == 8
    assert ttv_b200.plan(2, [2, (1 << 20) + 1, 128], [1, 2, 3], dtype="f32")["kernel"] == 2 and ttv_b200.plan(2, [3, (1 << 19) + 1, 64], [1, 2, 3], dtype="f64")["kernel"] == 2
    # a single huge fiber: split n_q across CTAs
    pl = ttv_b200.plan(1, [1 << 26, 2], [1, 2], dtype="f32")
    assert pl["kernel"] == 1 and pl["ksplit"] > 64 and pl["workspace_bytes"] == pl["ksplit"] * 2 * 4
    # complex<double>: one element per 16-byte load
    assert ttv_b200.plan(2, [64, 64, 64], [1, 2, 3], dtype="c128")["vec"] == 1
    # forcing
    assert ttv_b200.plan(1, [64, 64], [1, 2], dtype="f32", kernel="col")["kernel"] == 2
    with pytest.raises(ttv_b200.TTVError):
        ttv_b200.plan(2, [64, 64, 64], [1, 2, 3], dtype="f32", kernel="dot")
    assert ttv_b200.plan(2, [512, 512, 512], [1, 2, 3], dtype="f32", ksplit=3)["ksplit"] == 3
    assert ttv_b200.plan(2, [512, 512, 512], [1, 2, 3], dtype="f32", flags=4)["vec"] == 1


def test_no_cpu_fallback_without_device():
    if ttv_b200.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(ttv_b200.TTVError) as e:
        ttv_b200.ttv(2, np.ones((3, 4)), np.ones(4))
    assert e.value.status == 40 and "no CPU fallback" in str(e.value)


def test_chain_plan_orders():
    from ttv_b200.ttvpy import chain_plan
    # backward: highest mode first, numbering of the lower modes is unaffected
    assert chain_plan(2, (3, 2, 4, 5), "backward") == [(4, 2), (3, 1), (1, 0)]
    # forward: always mode 1 until q is reached, then always mode 2
    assert chain_plan(2, (3, 2, 4, 5), "forward") == [(1, 0), (2, 1), (2, 2)]
    assert chain_plan(1, (3, 2, 4, 5), "forward") == [(2, 0), (2, 1), (2, 2)]
    # optimal: longest vector first
    assert chain_plan(1, (3, 2, 4, 5), "optimal") == [(4, 2), (3, 1), (2, 0)]
    assert chain_plan(4, (3, 9, 4, 5), "optimal") == [(2, 1), (2, 2), (1, 0)]


def test_native_chain_plan_matches_the_restated_schedule():
    """ttv_b200_chain_plan (the schedule ttv_b200_ttvs follows) against the Python restatement of wrapped_ttv.cpp:135-192
    on random shapes, every q and order; and against the reference's own loops for backward / forward"""
    import random
    from ttv_b200 import api
    from ttv_b200.ttvpy import chain_plan
    rng = random.Random(7)
    for _ in range(300):
        p = rng.randint(2, 9)
        shape = [rng.choice([1, 2, 3, 5, 5, 8, 13]) for _ in range(p)]
        for q in range(1, p + 1):
            for order in ("optimal", "backward", "forward"):
                got = api.chain_plan(q, shape, order)
                assert got == chain_plan(q, shape, order), (shape, q, order)
                assert sorted(j for _, j in got) == list(range(p - 1))
            # wrapped_ttv.cpp:135-146 (backward) and :147-156 (forward), written out the way the reference loops
            back = [(p - 1 if q == p else p, p - 2)]
            r0 = back[0][0]
            back += [(r1, r1 - 2) for r1 in range(r0 - 1, q, -1)]
            back += [(r2, r2 - 1) for r2 in range((r0 - 1) if q == p else (q - 1), 0, -1)]
            assert api.chain_plan(q, shape, "backward") == back, (shape, q)
            fwd = [(2 if q == 1 else 1, 0)] + [(1, r1 - 1) for r1 in range(2, q)] + [(2, r2 - 1) for r2 in range(2 if q == 1 else q, p)]
            assert api.chain_plan(q, shape, "forward") == fwd, (shape, q)
    with pytest.raises(ttv_b200.TTVError) as e:
        api.chain_plan(0, (2, 3), "optimal")
    assert e.value.status == 2


def test_native_chain_needs_a_device():
    if ttv_b200.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    from ttv_b200 import ttvpy
    with pytest.raises(ttv_b200.TTVError) as e:
        ttvpy.ttvs(1, np.ones((3, 4, 2)), [np.ones(4), np.ones(2)])
    assert e.value.status == 40 and "no CPU fallback" in str(e.value)


def test_chooser_only_picks_instantiated_kernels():
    """fuzz of the canonical view: the (nu, ku) batch shape must be one the dispatcher instantiates (launch.cu), the
    thread tile must fit the CTA, vectors must divide the extent they run along"""
    import math
    import random
    rng = random.Random(1)
    size = {"f32": 4, "f64": 8, "c64": 8, "c128": 16, "i32": 4, "i64": 8}
    for dt in size:
        for _ in range(1500):
            outer = rng.choice([1, 2, 3, 7, 100, 5000, 10 ** 6, 10 ** 8])
            nq = rng.choice([1, 2, 3, 4, 5, 8, 16, 21, 23, 64, 215, 512, 1625, 4096, 65536, 10 ** 6])
            inner = rng.choice([1, 1, 2, 3, 4, 6, 16, 21, 23, 64, 100, 512, 1625, 65536, 10 ** 6])
            pl = ttv_b200.plan_view(outer, nq, inner, dtype=dt)
            wide = pl["vec"] * size[dt] >= 16
            key = (pl["nu"], pl["ku"])
            if pl["kernel"] == 3:       # STREAM: slabs through shared memory, one kernel shape
                assert nq * inner * size[dt] <= 75 * 1024 and nq * size[dt] <= 8192 and pl["smem_bytes"] <= 227 * 1024 and pl["ksplit"] == 1
                assert pl["ctas"] <= 148 * (2 if pl["smem_bytes"] + 1024 <= 227 * 1024 // 2 else 1)      # what an SM can hold
                continue
            if pl["kernel"] == 5:       # DOTF: short aligned fibers as one flat stream
                vec = 16 // size[dt]
                assert inner == 1 and nq % vec == 0 and nq // vec <= (128 if size[dt] == 16 else 48) and pl["vec"] == vec and pl["ksplit"] == 1
                assert pl["smem_bytes"] == (nq + 2048) * size[dt]
                continue
            if pl["kernel"] == 9:       # DOTP: fibers of two 4- / 8-byte elements, vectors of whole fibers, nothing shared
                assert inner == 1 and nq == 2 and size[dt] <= 8 and pl["vec"] == 16 // size[dt] and pl["ku"] in (4, 8) and pl["smem_bytes"] == 0 and pl["ksplit"] == 1
                assert pl["ctas"] == max(1, -(-(outer * 2 * size[dt] // 16) // (256 * pl["ku"])))          # one tile per CTA
                continue
            if pl["kernel"] == 10 and size[dt] == 4 and inner == 2 and nq <= 32 and nq & (nq - 1) == 0:
                # COLF, tiny slabs of two-element rows: consecutive lanes on consecutive vectors, transposing butterfly
                G = nq // 2
                assert (pl["tx"], pl["ty"], pl["to"], pl["nu"], pl["ku"]) == (1, G, 2, 32 // G, 8) and pl["smem_bytes"] == 0 and pl["ksplit"] == 1
                assert pl["ctas"] == -(-(-(-outer // (8 * (32 // G)))) // 8) and pl["workspace_bytes"] == 0          # a CTA per eight items
                continue
            if pl["kernel"] == 10:      # COLF: narrow / odd rows as a flat stream of super-rows, a warp per slab (partition) or several short slabs per warp
                vec = 16 // size[dt]
                g = math.gcd(inner, vec)
                L, R = inner // g, vec // g
                assert inner > 1 and g < vec and L <= 32 and (nq % R == 0 or outer == 1) and nq * inner * size[dt] >= 128 and (size[dt] == 4 or L <= 8)
                assert (pl["tx"], pl["to"]) == (L, R) and 1 <= pl["ty"] <= 32 // L and pl["nu"] * pl["ty"] * L <= 32 and pl["smem_bytes"] == 4096
                assert pl["nu"] == 1 or (pl["ksplit"] == 1 and pl["ty"] <= max(1, nq // R // 8))         # several slabs per warp: short slabs only
                groups = -(-outer // pl["nu"])
                cap = 148 * (4 if size[dt] == 8 else 6)
                if size[dt] == 4 and inner == 2 and pl["ksplit"] == 1 and nq // 2 <= 8 * pl["ty"]:
                    cap = 148 * 32                                      # the short-slab form of two-element rows: more, shorter-lived CTAs
                assert pl["ksplit"] >= 1 and pl["ctas"] == min(-(-groups * pl["ksplit"] // 8), cap)
                assert pl["workspace_bytes"] == (pl["ksplit"] > 1) * pl["ksplit"] * outer * inner * size[dt]
                continue
            if pl["kernel"] == 8:       # STREAMK: rows of a few elements under a long contraction, staged through shared memory
                assert 1 < inner <= 16 and size[dt] == 4 and inner % 2 == 1 and nq >= 4096              # odd rows of 4-byte elements
                assert outer > 1 and nq % ((16 // size[dt]) // math.gcd(inner, 16 // size[dt])) != 0      # what COLF cannot take
                assert pl["smem_bytes"] <= 113 * 1024 and pl["ctas"] <= 148 * 2 and pl["ksplit"] >= 1 and pl["threads"] == 256
                continue
            if pl["kernel"] == 4:       # COLX: odd wide rows, 16-byte loads at any phase
                assert (key in {(1, 8), (2, 4), (4, 2)} and wide or key == (8, 2) and size[dt] == 4) and size[dt] < 16, (dt, outer, nq, inner, pl)
                assert inner * size[dt] >= 2048 and inner % (16 // size[dt]) != 0
                if -(-nq // (16 // size[dt])) < 48:      # short contraction: warp-autonomous form, no shared memory
                    # (complex<float> takes the realigned form COLR, whose outputs leave through a warp-private shared-memory strip)
                    assert (pl["tx"], pl["ty"]) == (32, 1) and (pl["smem_bytes"] == 0 or dt == "c64") and pl["ku"] % pl["vec"] == 0 and pl["vec"] == 2
                else:
                    assert pl["ty"] == 16 // size[dt] and pl["tx"] * pl["ty"] <= 256 and pl["smem_bytes"] <= 100 * 1024
                continue
            if pl["kernel"] == 1:
                allowed = {(1, 8), (2, 4), (4, 2), (8, 1)} if wide else {(1, 16), (2, 8), (4, 4), (8, 2), (8, 1)}
            else:
                allowed = {(1, 8), (2, 4), (4, 2)} if wide else {(1, 16), (2, 8), (4, 4), (8, 2)}
                if wide and pl["ksplit"] == 1 and pl["ctas"] < 148 * 10:
                    allowed = allowed | {(1, 16)}       # few CTAs: 16 vector loads per batch, 2 CTAs per SM
            if pl["ty"] > 1:
                allowed = allowed | {(1, 8)}        # b read directly from L2: the one batch shape of that variant
            assert key in allowed, (dt, outer, nq, inner, pl)
            assert pl["tx"] * pl["ty"] * pl["to"] <= pl["threads"] <= 256
            assert pl["ksplit"] >= 1 and pl["ctas"] >= 1 and pl["smem_bytes"] <= 100 * 1024
            if pl["kernel"] == 2:
                assert inner % pl["vec"] == 0


def test_layout_and_strides_of_numpy_arrays():
    """the front end reads arrays in place: layout = order of the strides, wa = the strides, and the honour-strides
    flag only when they are not the packed strides of that layout"""
    from ttv_b200.api import _layout_of
    x = np.zeros((4, 5, 6))
    assert _layout_of(x, None) == ([3, 2, 1], [30, 6, 1], False)                       # C order = last-order
    assert _layout_of(np.asfortranarray(x), None) == ([1, 2, 3], [1, 4, 20], False)    # F order = first-order
    assert _layout_of(x.transpose(1, 0, 2), None) == ([3, 1, 2], [6, 30, 1], False)    # a permuted, still packed layout
    assert _layout_of(x[:, :3, :], None) == ([3, 2, 1], [30, 6, 1], True)              # slice: padded
    assert _layout_of(x[:, :, ::2], None) == ([3, 2, 1], [30, 6, 2], True)             # fastest mode strided
    assert _layout_of(x[::-1], None) is None                                           # reversed axis: copy
    assert _layout_of(np.broadcast_to(np.zeros(6), (4, 5, 6)), None) is None           # broadcast axes: copy
    pia, wa, honor = _layout_of(np.zeros((4, 1, 6)), None)
    assert ttv_b200.is_valid_strides(pia, wa) and not honor


def test_bench_stdout_carries_only_the_json_line():
    """bench.py's contract is ONE JSON line on stdout; libraries that write to fd 1 behind Python's back (NCCL's version
    banner did at N = 2) must end up on stderr"""
    import subprocess
    import sys
    code = (
        "import sys, ctypes\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import bench\n"
        "bench._guard_stdout()\n"
        "libc = ctypes.CDLL(None)\n"
        "libc.puts(b'NCCL version x.y.z')\n"
        "libc.fflush(None)\n"
        "print('a python print')\n"
        "bench.emit({'metric': 'm', 'value': 1.5})\n"
    )
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout == '{"metric": "m", "value": 1.5}\n', r.stdout
    assert "NCCL version x.y.z" in r.stderr and "a python print" in r.stderr


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's own CPU implementation, oracle/_ref or the oracle port) needs no GPU:
    one JSON line with the keys the driver reads.  Needs ~17 GiB of host memory for the 256^4 fp32 tensor."""
    import json
    import subprocess
    import sys
    import psutil
    if psutil.virtual_memory().available < 24 * 2 ** 30:
        pytest.skip("not enough free host memory for the 16 GiB workload")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:2000]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "TTV effective HBM GB/s" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "256" in d["config"]["workload"] and "model" not in d["config"]


def test_bench_reference_arm_under_torchrun_prints_once():
    """launched the way the driver launches N > 1 (one rank per GPU): rank 0 alone runs the CPU arm and prints the line,
    the other rank exits 0 without work"""
    import json
    import subprocess
    import sys
    import psutil
    if psutil.virtual_memory().available < 24 * 2 ** 30:
        pytest.skip("not enough free host memory for the 16 GiB workload")
    import socket
    with socket.socket() as sock:                       # a port nobody is listening on right now
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1, r.stdout[:2000]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0


def test_bench_own_arm_fails_loudly_without_a_gpu():
    """no CPU fallback: without a CUDA device the product arm of bench.py must refuse to run, not print a number"""
    import subprocess
    import sys
    if ttv_b200.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout) and not any(l.startswith("{") for l in r.stdout.splitlines())
