"""Parity of the CUDA path with the oracle, through the C-ABI.  Needs a B200 (pytest -m gpu).

Bit-exact for integer element types and for integer-valued float data; within 2*n_q*eps*sum|a||b| per element for
float/double/complex data in [-1,1).  The test grid of the reference (gtest_tlib_ttv.cpp:192-425) is run
in full for double like the reference does, plus every other element type on a thinned grid, plus what the reference
never tests: non-power-of-two extents, extents of 1, order > 4, forced kernels / split n_q / scalar loads, device
pointers, accumulate, large sizes through size-independent properties."""
from __future__ import annotations

import os

import numpy as np
import pytest

import ttv_b200
from conftest import (all_layouts, assert_close, fold, random_case, real_case, reference_expected, reference_init,
                      reference_shapes)

pytestmark = pytest.mark.gpu

ALL_DTYPES = [np.float32, np.float64, np.complex64, np.complex128, np.int32, np.int64]
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ttv_golden.npz")


def run_lowlevel(q, a, na, pia, b, *, c0=None, **opts):
    """call the C-like interface with host buffers the way gtest_tlib_ttv.cpp:96-135 does"""
    p = len(na)
    nc = ttv_b200.generate_output_shape(na, q)
    pic = ttv_b200.generate_output_layout(pia, q)
    wa = ttv_b200.generate_strides(na, pia)
    wc = ttv_b200.generate_strides(nc, pic)
    c = np.zeros(int(np.prod(nc, dtype=object)), a.dtype) if c0 is None else c0
    ttv_b200.ttv_lowlevel(q, p, a, na, wa, pia, b, [len(b)], c, nc, wc, pic, **opts)
    return c


# ---- the reference's own grid -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("order", [2, 3, 4])
def test_reference_grid_double(order):
    """TEST(TensorTimesVector, *): double, {2,4,8}^p, all layouts, every q, b = 1, closed form :132"""
    before = ttv_b200.launch_count()
    n_calls = 0
    for na in reference_shapes(order):
        for pia in all_layouts(order):
            for q in range(1, order + 1):
                a = reference_init(na, pia, q, np.float64)
                b = np.ones(na[q - 1], np.float64)
                c = run_lowlevel(q, a, na, pia, b)
                assert np.array_equal(c, reference_expected(na, q, c.size, np.float64)), (na, pia, q)
                n_calls += 1
    assert ttv_b200.launch_count() - before >= n_calls      # the CUDA kernels ran, nothing else could have


@pytest.mark.parametrize("dtype", [np.float32, np.complex64, np.complex128, np.int32, np.int64])
def test_reference_grid_other_types(dtype, oracle):
    rng = np.random.default_rng(5)
    for order in (2, 3, 4):
        shapes = reference_shapes(order)
        for na in shapes[:: max(1, len(shapes) // 9)]:
            for pia in all_layouts(order):
                for q in range(1, order + 1):
                    a, b = random_case(rng, na, q, dtype)
                    c = run_lowlevel(q, a, na, pia, b)
                    assert np.array_equal(c, oracle.ttv(q, a, na, pia, b)), (na, pia, q, dtype)


@pytest.mark.parametrize("policy", [("seq", "slice", "none"), ("par_loop", "slice", "all"), ("par_loop", "slice", "outer"),
                                    ("par_blas", "subtensor", "all"), ("par_taskloop", "subtensor", "none")])
def test_policy_hints_select_the_same_path(policy, oracle):
    rng = np.random.default_rng(6)
    na, pia = (4, 8, 2, 4), (3, 1, 4, 2)
    for q in range(1, 5):
        a, b = random_case(rng, na, q, np.float64)
        c = run_lowlevel(q, a, na, pia, b, execution=policy[0], slicing=policy[1], fusion=policy[2])
        assert np.array_equal(c, oracle.ttv(q, a, na, pia, b, "slice" if policy[1] == "slice" else "subtensor"))


# ---- what the reference does not test ----------------------------------------------------------------------------------------
ODD_SHAPES = [(3, 5), (1, 7), (7, 1), (257, 3), (3, 5, 7), (1, 4, 1), (5, 1, 3), (33, 2, 17), (2, 129, 3), (3, 4, 5, 2),
              (1, 1, 6, 2), (2, 3, 1, 5, 2), (2, 2, 3, 2, 2, 3), (2, 2, 2, 2, 2, 2, 3), (2, 1, 2, 2, 1, 2, 2, 3)]


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_non_power_of_two_and_unit_extents(dtype, oracle):
    rng = np.random.default_rng(8)
    for na in ODD_SHAPES:
        p = len(na)
        layouts = all_layouts(p) if p <= 3 else [tuple(ttv_b200.generate_k_order_layout(p, k)) for k in (1, 0, 2)] + \
            [tuple(int(x) for x in rng.permutation(p) + 1) for _ in range(3)]
        for pia in layouts:
            for q in range(1, p + 1):
                a, b = random_case(rng, na, q, dtype)
                c = run_lowlevel(q, a, na, pia, b)
                assert np.array_equal(c, oracle.naive(q, a, na, pia, b)), (na, pia, q, dtype)


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_random_orders_shapes_layouts(dtype, oracle):
    """seeded fuzz over what the reference's grid never varies together: order 2..9, extents that are odd / prime / 1 /
    one or two large ones (so that every kernel family and its ragged edges are reached through the chooser), random
    non-hierarchical layouts, every q; device and host pointers alternate.  Bit-exact against the oracle's loop nest."""
    import torch
    rng = np.random.default_rng(20261017 + ALL_DTYPES.index(dtype))
    small = [1, 2, 3, 4, 5, 7, 8, 9, 16, 21, 23]
    large = [31, 40, 64, 84, 127, 215, 256, 333, 512, 1025, 1625, 4099]
    seen = set()
    for case in range(60):
        p = int(rng.integers(2, 10))
        na = [int(rng.choice(small)) for _ in range(p)]
        for _ in range(int(rng.integers(0, 3))):                      # up to two large modes
            na[int(rng.integers(0, p))] = int(rng.choice(large))
        while int(np.prod(na, dtype=object)) > 3_000_000:              # keep the oracle fast: shrink the largest mode
            na[int(np.argmax(na))] = max(1, max(na) // 2)
        pia = [int(x) + 1 for x in rng.permutation(p)]
        for q in range(1, p + 1):
            a, b = random_case(rng, na, q, dtype)
            want = oracle.ttv(q, a, na, pia, b)
            if (case + q) % 2:
                got = run_lowlevel(q, a, na, pia, b)
            else:
                ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
                tc = torch.full((want.size,), 7, dtype=ta.dtype, device="cuda")          # C is overwritten
                nc = ttv_b200.generate_output_shape(na, q); pic = ttv_b200.generate_output_layout(pia, q)
                ttv_b200.ttv_lowlevel(q, p, ta, na, ttv_b200.generate_strides(na, pia), pia, tb, [len(b)], tc, nc,
                                      ttv_b200.generate_strides(nc, pic), pic)
                torch.cuda.synchronize()
                got = tc.cpu().numpy()
            assert np.array_equal(got, want), (na, pia, q)
            seen.add(ttv_b200.plan(q, na, pia, dtype=ttv_b200.api.dtype_code(a))["kernel"])
    assert {1, 2}.issubset(seen), seen                                 # at least the two main families were exercised


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
def test_rounding_within_stated_tolerance(dtype, oracle):
    rng = np.random.default_rng(9)
    for na, pia in [((64, 300, 5), (1, 2, 3)), ((64, 300, 5), (3, 2, 1)), ((5, 7, 1000), (2, 3, 1)), ((1000, 37), (1, 2)),
                    ((1000, 37), (2, 1)), ((12, 20, 6, 9), (4, 2, 1, 3))]:
        for q in range(1, len(na) + 1):
            a, b = real_case(rng, na, q, dtype)
            wide, mag = oracle.naive(q, a, na, pia, b, want_abs=True)
            ref = oracle.ttv(q, a, na, pia, b)
            c = run_lowlevel(q, a, na, pia, b)
            assert_close(c, wide, mag, na[q - 1], dtype, f"cuda vs wide {na} {pia} q={q}")
            assert_close(c, ref, 2 * mag, na[q - 1], dtype, f"cuda vs oracle {na} {pia} q={q}")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.int32])
@pytest.mark.parametrize("variant", [dict(ksplit=1), dict(ksplit=3), dict(ksplit=7), dict(flags=4), dict(flags=4, ksplit=2),
                                     dict(kernel="col"), dict(kernel="col", ksplit=4)])
def test_forced_variants(dtype, variant, oracle):
    """every kernel family / split / vector width on shapes that reach each of their code paths"""
    rng = np.random.default_rng(10)
    shapes = [((8, 64, 6), (1, 2, 3)), ((300, 40), (1, 2)), ((300, 40), (2, 1)), ((2, 500, 3), (1, 2, 3)), ((16, 9, 33), (3, 1, 2)),
              ((1024, 70), (1, 2)), ((4, 4100), (2, 1))]
    for na, pia in shapes:
        for q in range(1, len(na) + 1):
            a, b = random_case(rng, na, q, dtype)
            c = run_lowlevel(q, a, na, pia, b, **variant)
            assert np.array_equal(c, oracle.ttv(q, a, na, pia, b)), (na, pia, q, variant)


def test_golden_fixtures_of_the_reference(oracle):
    g = np.load(GOLDEN, allow_pickle=False)
    for i in range(int(g["count"])):
        na = g[f"na_{i}"].tolist(); pia = g[f"pia_{i}"].tolist(); q = int(g[f"q_{i}"])
        a, b, c_ref = g[f"a_{i}"], g[f"b_{i}"], g[f"c_{i}"]
        c = run_lowlevel(q, a, na, pia, b)
        if bool(g[f"exact_{i}"]):
            assert np.array_equal(c, c_ref), (i, na, pia, q, a.dtype)
        else:
            _, mag = oracle.naive(q, a, na, pia, b, want_abs=True)
            assert_close(c, c_ref, mag, na[q - 1], a.dtype, what=f"golden case {i}")


def test_known_answers():
    a = np.arange(1, 25, dtype=np.float32)        # example/interface3.cpp
    assert run_lowlevel(2, a, (4, 3, 2), (1, 2, 3), np.ones(3, np.float32)).tolist() == [15, 18, 21, 24, 51, 54, 57, 60]
    A = np.arange(24, dtype=np.float64).reshape(3, 2, 4)     # ttvpy/README.md:61-66
    assert ttv_b200.ttvpy.ttv(1, A, np.arange(3, dtype=np.float64)).tolist() == [[40, 43, 46, 49], [52, 55, 58, 61]]


def test_overwrite_and_accumulate(oracle):
    """C is overwritten by default (what the BLAS build of the reference does, matrix_times_vector.h:213-215);
    FLAG_ACCUMULATE gives the non-BLAS column kernel's C += (matrix_times_vector.h:124)."""
    rng = np.random.default_rng(12)
    for na, pia in [((4, 3, 2), (1, 2, 3)), ((40, 30), (1, 2)), ((6, 50, 7), (2, 3, 1))]:
        for q in range(1, len(na) + 1):
            a, b = random_case(rng, na, q, np.float64)
            expect = oracle.ttv(q, a, na, pia, b)
            c0 = np.full(expect.size, 100.0)
            assert np.array_equal(run_lowlevel(q, a, na, pia, b, c0=c0.copy()), expect)
            assert np.array_equal(run_lowlevel(q, a, na, pia, b, c0=c0.copy(), flags=1), expect + 100.0)
            assert np.array_equal(run_lowlevel(q, a, na, pia, b, c0=c0.copy(), flags=1, ksplit=3), expect + 100.0)


def test_errors_surface_with_reference_messages():
    a = np.zeros(24); b = np.zeros(3); c = np.zeros(8)
    with pytest.raises(ttv_b200.TTVError) as e:
        ttv_b200.ttv_lowlevel(2, 3, a, [4, 3, 2], [1, 4, 12], [1, 2, 3], b, [4], c, [4, 2], [1, 4], [1, 2])
    assert e.value.status == 13 and str(e.value).startswith("Error in tlib::tensor_times_vector: contraction dimension")
    with pytest.raises(ttv_b200.TTVError) as e:
        ttv_b200.ttv_lowlevel(2, 3, a, [4, 3, 2], [1, 4, 12], [1, 2, 3], b, [3], c, [4, 2], [2, 1], [2, 1])
    assert e.value.status == 20


# ---- device pointers ------------------------------------------------------------------------------------------------------
def test_device_pointers_in_place(oracle):
    import torch
    rng = np.random.default_rng(13)
    for dtype in ALL_DTYPES:
        for na, pia in [((6, 50, 7), (2, 3, 1)), ((129, 65), (1, 2))]:
            for q in range(1, len(na) + 1):
                a, b = random_case(rng, na, q, dtype)
                ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
                nc = ttv_b200.generate_output_shape(na, q); pic = ttv_b200.generate_output_layout(pia, q)
                tc = torch.full((int(np.prod(nc)),), 7, dtype=ta.dtype, device="cuda")
                ttv_b200.ttv_lowlevel(q, len(na), ta, na, ttv_b200.generate_strides(na, pia), pia, tb, [len(b)], tc, nc,
                                      ttv_b200.generate_strides(nc, pic), pic)
                assert np.array_equal(tc.cpu().numpy(), oracle.ttv(q, a, na, pia, b))
    # mixing host and device buffers is an error, not a silent copy
    with pytest.raises(ttv_b200.TTVError) as e:
        ttv_b200.ttv_lowlevel(1, 2, ta, [129, 65], [1, 129], [1, 2], b if len(b) == 129 else np.zeros(129, a.dtype), [129],
                              tc[:65], [65], [1], [1])
    assert e.value.status == 41


def test_tensor_level_interface_numpy_and_torch(oracle):
    import torch
    rng = np.random.default_rng(14)
    A = rng.integers(-5, 6, (5, 4, 6, 3)).astype(np.float64)
    subs = {1: "ijkl,i->jkl", 2: "ijkl,j->ikl", 3: "ijkl,k->ijl", 4: "ijkl,l->ijk"}
    for q in range(1, 5):
        b = rng.integers(-5, 6, A.shape[q - 1]).astype(np.float64)
        want = np.einsum(subs[q], A, b)
        assert np.array_equal(ttv_b200.ttv(q, A, b), want)                                   # C order  (last-order)
        assert np.array_equal(ttv_b200.ttv(q, np.asfortranarray(A), b), want)                # F order  (first-order)
        got = ttv_b200.ttv(q, torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda())
        assert np.array_equal(got.cpu().numpy(), want)


# ---- sizes the oracle cannot hold: size-independent properties ------------------------------------------------------------------
def test_large_shapes_by_properties(oracle):
    """BASELINE config 1 at full size (512^3 fp32, q=2) and a 2^31+-element index range: linearity in b, agreement of
    independent kernel variants, and sampled fibers against a host long-double dot on generator-defined data."""
    import torch
    n = 512
    seed_a, seed_b = 0x77170001, 0x77170002
    A = torch.empty(n * n * n, dtype=torch.float32, device="cuda")
    ttv_b200.fill(A, seed_a)
    b1 = torch.empty(n, dtype=torch.float32, device="cuda"); ttv_b200.fill(b1, seed_b)
    b2 = torch.empty(n, dtype=torch.float32, device="cuda"); ttv_b200.fill(b2, seed_b + 5)
    a_host_sample = oracle.fill("f32", 4096, seed_a)
    assert np.array_equal(A[:4096].cpu().numpy(), a_host_sample)            # device generator == oracle generator
    na, pia = [n, n, n], [1, 2, 3]
    wa = ttv_b200.generate_strides(na, pia)
    bh = b1.cpu().numpy().astype(np.longdouble)
    rng = np.random.default_rng(15)
    for q in (1, 2, 3):
        nc = ttv_b200.generate_output_shape(na, q); pic = ttv_b200.generate_output_layout(pia, q)
        wc = ttv_b200.generate_strides(nc, pic)
        outs = []
        for bb, kw in ((b1, {}), (b2, {}), (b1 + b2, {}), (b1, dict(ksplit=4)), (b1, dict(flags=4))):
            c = torch.empty(n * n, dtype=torch.float32, device="cuda")
            ttv_b200.ttv_lowlevel(q, 3, A, na, wa, pia, bb, [n], c, nc, wc, pic, **kw)
            outs.append(c.cpu().numpy().astype(np.float64))
        c1, c2, c12, c1_split, c1_scalar = outs
        tol = 8 * n * np.finfo(np.float32).eps
        assert np.abs(c12 - (c1 + c2)).max() <= tol * 4                     # linearity in b
        assert np.abs(c1 - c1_split).max() <= tol and np.abs(c1 - c1_scalar).max() <= tol
        # sampled fibers vs a host long-double dot on regenerated data (SURVEY 8c "too big for host")
        outer, nq, inner = fold(na, pia, q)
        for j in rng.integers(0, n * n, 64):
            o, i = divmod(int(j), inner)
            idx = (o * nq + np.arange(nq)) * inner + i
            fiber = np.array([oracle.fill("f32", 1, seed_a, first=int(e))[0] for e in idx], dtype=np.longdouble)
            want = float(np.dot(fiber, bh))
            assert abs(c1[j] - want) <= 2 * nq * (np.finfo(np.float32).eps / 2) * float(np.dot(np.abs(fiber), np.abs(bh)))


def test_more_than_2_to_32_elements():
    """64-bit indexing: 2^32 + 2^20 int32 elements (16.4 GiB), q in every position; checksums in closed form."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < 40 * 2 ** 30:
        pytest.skip("needs 40 GiB of free HBM")
    na = [1024, 4097, 1024]                 # 2^32 + 2^20 elements
    total = na[0] * na[1] * na[2]
    A = torch.ones(total, dtype=torch.int32, device="cuda")
    # A = 1 everywhere except one marked element near the end of the index range
    marked = total - 12345
    A[marked] = 1000
    pia = [1, 2, 3]
    wa = ttv_b200.generate_strides(na, pia)
    for q in (1, 2, 3):
        nq = na[q - 1]
        b = torch.arange(1, nq + 1, dtype=torch.int32, device="cuda")
        nc = ttv_b200.generate_output_shape(na, q); pic = ttv_b200.generate_output_layout(pia, q)
        c = torch.empty(total // nq, dtype=torch.int32, device="cuda")
        ttv_b200.ttv_lowlevel(q, 3, A, na, wa, pia, b, [nq], c, nc, ttv_b200.generate_strides(nc, pic), pic)
        base = nq * (nq + 1) // 2
        outer, _, inner = fold(na, pia, q)
        o, rem = divmod(marked, nq * inner)
        k, i = divmod(rem, inner)
        expect_marked = base + 999 * (k + 1)
        j = o * inner + i
        assert int(c[j]) == expect_marked
        assert int((c != base).sum()) == 1


def test_multi_products_on_one_tensor(oracle):
    """ttv_b200_multi: all modes of one tensor in one call (A staged once); identical to separate calls"""
    import torch
    rng = np.random.default_rng(16)
    for dtype in (np.float64, np.int32, np.complex64):
        na, pia = (6, 9, 4, 5), (2, 4, 1, 3)
        a, _ = random_case(rng, na, 1, dtype)
        qs = [1, 2, 3, 4, 2]
        bs = [random_case(rng, na, q, dtype)[1] for q in qs]
        want = [oracle.ttv(q, a, na, pia, b) for q, b in zip(qs, bs)]
        got = ttv_b200.ttv_multi(qs, a, na, pia, bs)                                  # host buffers
        assert all(np.array_equal(g, w) for g, w in zip(got, want))
        ta = torch.from_numpy(a).cuda()
        got = ttv_b200.ttv_multi(qs, ta, na, pia, [torch.from_numpy(b).cuda() for b in bs])   # device buffers
        assert all(np.array_equal(g.cpu().numpy(), w) for g, w in zip(got, want))
    with pytest.raises(ttv_b200.TTVError) as e:
        ttv_b200.ttv_multi([1, 5], a, na, pia, [bs[0], bs[0]], cs=[np.empty(8, a.dtype), np.empty(8, a.dtype)])
    assert e.value.status == 2


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_stream_kernel_small_slabs(dtype, oracle):
    """kernel="stream": slabs staged through shared memory by TMA bulk copies; odd sizes, many chunks, ragged last
    chunk, array ends that are not a multiple of 16 bytes, accumulate"""
    rng = np.random.default_rng(17)
    cases = [((23, 23, 301), (1, 2, 3), 2), ((23, 1013), (1, 2), 1), ((21, 21, 77), (1, 2, 3), 1), ((3, 2, 5000), (1, 2, 3), 2),
             ((5, 7, 3, 211), (2, 1, 3, 4), 1), ((5, 7, 3, 211), (2, 1, 3, 4), 2), ((9, 4099), (1, 2), 1), ((2, 100003), (1, 2), 1),
             ((1, 7, 13), (1, 2, 3), 2), ((40, 5, 6000), (1, 2, 3), 1),
             # fibers of even length are walked skewed (bank conflicts otherwise)
             ((84, 3001), (1, 2), 1), ((4, 70001), (1, 2), 1), ((24, 5003), (1, 2), 1), ((128, 1001), (1, 2), 1), ((2, 3, 4099), (1, 2, 3), 1)]
    # slabs that get a stage of their own (36 .. 75 KB): odd rows, ragged last chunk, a slab that ends the tensor unaligned
    cases += [((529, 23, 9), (1, 2, 3), 2), ((73, 73, 11), (1, 2, 3), 2), ((441, 21, 5), (1, 2, 3), 2), ((2, 4500, 7), (1, 2, 3), 1),
              ((1201, 9, 3), (1, 2, 3), 2)]
    for na, pia, q in cases:
        a, b = random_case(rng, na, q, dtype)
        want = oracle.ttv(q, a, na, pia, b)
        name = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64", np.dtype(np.complex64): "c64",
                np.dtype(np.complex128): "c128", np.dtype(np.int32): "i32", np.dtype(np.int64): "i64"}[np.dtype(dtype)]
        try:
            pl = ttv_b200.plan(q, na, pia, dtype=name, kernel="stream")
        except ttv_b200.TTVError as exc:          # slab larger than three stages of shared memory hold, or b too long
            k = list(pia).index(q)
            slab = int(np.prod([na[m - 1] for m in pia[: k + 1]])) * np.dtype(dtype).itemsize
            assert exc.status == 32 and (slab > 74 * 1024 or na[q - 1] * np.dtype(dtype).itemsize > 8192), (na, pia, q, slab)
            continue
        assert pl["kernel"] == 3
        c = run_lowlevel(q, a, na, pia, b, kernel="stream")
        assert np.array_equal(c, want), (na, pia, q, dtype)
        c0 = np.full(want.size, 3, dtype)
        assert np.array_equal(run_lowlevel(q, a, na, pia, b, c0=c0, kernel="stream", flags=1), want + 3)


@pytest.mark.parametrize("form", ["cta", "warp", "realign", "realign8", "auto"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.int32, np.int64])
def test_colx_kernel_unaligned_rows(dtype, form, oracle, monkeypatch):
    """kernel="colx", all forms (phase lanes across a CTA / all phases inside a warp / rows realigned at load time with
    shuffles, two batch depths): rows that start at every phase of a 16-byte line (inner % 4 = 1, 2, 3 and 0), several
    tiles per row, ragged last tile, a tensor whose last bytes are not a whole vector, n_q shorter than the phase lanes
    and not a multiple of the batch, split n_q, accumulate; and the shapes the chooser routes there by itself"""
    if form != "auto":
        monkeypatch.setenv("TTV_B200_COLX_WARP", {"cta": "0", "warp": "1", "realign": "2", "realign8": "2"}[form])
    if form == "realign8":
        monkeypatch.setenv("TTV_B200_KU", "8")
    rng = np.random.default_rng(12)
    name = {np.float32: "f32", np.float64: "f64", np.complex64: "c64", np.int32: "i32", np.int64: "i64"}[dtype]
    cases = [((37, 5, 3), (1, 2, 3), 2), ((1021, 9, 2), (1, 2, 3), 2), ((1021, 3), (1, 2), 2), ((333, 7, 5), (1, 2, 3), 3),
             ((3, 1033, 6), (2, 1, 3), 3), ((2055, 70, 3), (1, 2, 3), 2), ((5, 413, 1, 9), (1, 2, 3, 4), 4), ((64, 10, 3), (1, 2, 3), 2),
             ((2, 9, 3), (1, 2, 3), 2), ((4099, 33), (1, 2), 2), ((1022, 9, 3), (1, 2, 3), 2), ((6, 343, 5, 2), (1, 2, 3, 4), 3),
             ((250, 23, 7), (1, 2, 3), 2), ((127, 23, 9), (1, 2, 3), 2), ((23, 23, 23, 5), (1, 2, 3, 4), 4)]
    for na, pia, q in cases:
        a, b = random_case(rng, na, q, dtype)
        want = oracle.ttv(q, a, na, pia, b)
        for extra in (dict(), dict(ksplit=3)):
            pl = ttv_b200.plan(q, na, pia, dtype=name, kernel="colx", **extra)
            assert pl["kernel"] == 4
            c = run_lowlevel(q, a, na, pia, b, kernel="colx", **extra)
            assert np.array_equal(c, want), (na, pia, q, extra)
        c0 = np.full(want.size, 3, dtype)
        assert np.array_equal(run_lowlevel(q, a, na, pia, b, c0=c0, kernel="colx", flags=1), want + 3)
    # chosen automatically for wide odd rows
    for na, pia, q in [((1625, 40, 3), (1, 2, 3), 2), ((23, 23, 23, 7), (1, 2, 3, 4), 4)]:
        a, b = random_case(rng, na, q, dtype)
        if np.dtype(dtype).itemsize * int(np.prod(na[: q - 1])) >= 2048:
            assert ttv_b200.plan(q, na, pia, dtype=name)["kernel"] == 4
        assert np.array_equal(run_lowlevel(q, a, na, pia, b), oracle.ttv(q, a, na, pia, b)), (na, pia, q)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128, np.int32])
def test_b_read_directly_from_l2(dtype, oracle, monkeypatch):
    """lanes strung along n_q with a b that is too long to stay in shared memory read it inside the batches (no
    re-staging, no __syncthreads in the main loop): chosen automatically for long n_q, forced on the small shapes"""
    rng = np.random.default_rng(13)
    name = {np.float32: "f32", np.float64: "f64", np.complex128: "c128", np.int32: "i32"}[dtype]
    for na, pia, q in [((4, 6001, 3), (1, 2, 3), 2), ((70001, 3), (1, 2), 1), ((70004, 2), (1, 2), 1), ((2, 40000), (1, 2), 2)]:
        a, b = random_case(rng, na, q, dtype)
        want = oracle.ttv(q, a, na, pia, b)
        for extra in (dict(), dict(ksplit=3)):
            assert np.array_equal(run_lowlevel(q, a, na, pia, b, **extra), want), (na, pia, q, extra)
    monkeypatch.setenv("TTV_B200_BDIRECT", "1")
    for na, pia in [((8, 64, 6), (1, 2, 3)), ((300, 40), (1, 2)), ((300, 40), (2, 1)), ((2, 500, 3), (1, 2, 3)), ((16, 9, 33), (3, 1, 2)),
                    ((4, 4100), (2, 1))]:
        for q in range(1, len(na) + 1):
            a, b = random_case(rng, na, q, dtype)
            want = oracle.ttv(q, a, na, pia, b)
            for extra in (dict(), dict(ksplit=2), dict(flags=4)):
                assert np.array_equal(run_lowlevel(q, a, na, pia, b, **extra), want), (na, pia, q, extra)


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_dotf_kernel_short_fibers_as_one_stream(dtype, oracle):
    """kernel="dotf": fibers of 1 ... 256 vectors read as one flat stream, warp-private partial sums; ragged last chunk,
    fewer fibers than a chunk, odd and even vector counts (rotated summation), accumulate"""
    rng = np.random.default_rng(21)
    name = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64", np.dtype(np.complex64): "c64",
            np.dtype(np.complex128): "c128", np.dtype(np.int32): "i32", np.dtype(np.int64): "i64"}[np.dtype(dtype)]
    vec = 16 // np.dtype(dtype).itemsize
    for nv, fibers in [(1, 5000), (2, 3001), (3, 777), (5, 10000), (10, 2561), (21, 1300), (24, 999), (40, 650), (64, 300), (100, 70),
                       (255, 40), (256, 33), (7, 3), (12, 1)]:
        na, pia, q = (nv * vec, fibers), (1, 2), 1
        a, b = random_case(rng, na, q, dtype)
        want = oracle.ttv(q, a, na, pia, b)
        assert ttv_b200.plan(q, na, pia, dtype=name, kernel="dotf")["kernel"] == 5
        assert np.array_equal(run_lowlevel(q, a, na, pia, b, kernel="dotf"), want), (na, dtype)
        c0 = np.full(want.size, 3, dtype)
        assert np.array_equal(run_lowlevel(q, a, na, pia, b, c0=c0, kernel="dotf", flags=1), want + 3)
    # order 3, mode q first in the layout but not mode 1
    na, pia, q = (6, 4 * vec, 50), (2, 1, 3), 2
    a, b = random_case(rng, na, q, dtype)
    assert np.array_equal(run_lowlevel(q, a, na, pia, b, kernel="dotf"), oracle.ttv(q, a, na, pia, b))
    with pytest.raises(ttv_b200.TTVError):
        ttv_b200.plan(1, (vec + 1, 9), (1, 2), dtype=name, kernel="dotf") if vec > 1 else ttv_b200.plan(2, (4, 9), (1, 2), dtype=name, kernel="dotf")


def _padded(rng, na, wa, dtype):
    """a logical tensor X of shape na and the flat buffer that holds it with element strides wa (padding = 77)"""
    x = rng.integers(-8, 9, na).astype(dtype)
    span = 1 + sum((n - 1) * w for n, w in zip(na, wa))
    buf = np.full(span, 77, dtype)
    view = np.lib.stride_tricks.as_strided(buf, shape=na, strides=[w * buf.itemsize for w in wa])
    view[...] = x
    return x, buf


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128, np.int32])
def test_non_packed_strides(dtype, oracle):
    """padded strides in case 8: honoured like the reference's slice variants
    (tensor_times_vector.h:189-216), results against the oracle's slice nest on the same padded buffers and against the
    definition; the padding of C comes back untouched; host and device buffers"""
    import torch
    rng = np.random.default_rng(31)
    code = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.complex128): 3, np.dtype(np.int32): 4}[np.dtype(dtype)]
    cases = [
        # na, pia, q, wa, wc   (pic is derived; strides are valid = non-decreasing along the layout)
        ((4, 3, 2), (1, 2, 3), 2, (1, 8, 24), (1, 4)),             # padded leading dimension of A
        ((4, 3, 2), (1, 2, 3), 2, (1, 4, 12), (1, 6)),             # padded C only
        ((5, 6, 7, 3), (1, 2, 3, 4), 3, (1, 8, 50, 400), (1, 7, 45)),
        ((5, 6, 7, 3), (2, 1, 4, 3), 4, (9, 1, 200, 60), (7, 1, 50)),
        ((33, 5, 40, 2), (1, 2, 3, 4), 2, (1, 40, 200, 8000), (1, 40, 1600)),      # folds partly: modes 3, 4 stay packed
        ((3, 300, 4), (1, 2, 3), 2, (1, 4, 1300), (1, 5)),
    ]
    for na, pia, q, wa, wc in cases:
        p = len(na)
        nc = ttv_b200.generate_output_shape(na, q); pic = ttv_b200.generate_output_layout(pia, q)
        x, a = _padded(rng, na, wa, dtype)
        b = rng.integers(-8, 9, na[q - 1]).astype(dtype)
        want = np.tensordot(x, b, axes=([q - 1], [0]))
        span_c = 1 + sum((n - 1) * w for n, w in zip(nc, wc))
        assert ttv_b200.plan(q, na, pia, dtype=code, wa=list(wa), wc=list(wc))["kernel"] == 6
        for where in ("host", "device"):
            c = np.full(span_c, 55, dtype)
            if where == "host":
                ttv_b200.ttv_lowlevel(q, p, a, na, wa, pia, b, [len(b)], c, nc, wc, pic)
            else:
                ta, tb, tc = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(c).cuda()
                ttv_b200.ttv_lowlevel(q, p, ta, na, wa, pia, tb, [len(b)], tc, nc, wc, pic)
                c = tc.cpu().numpy()
            got = np.lib.stride_tricks.as_strided(c, shape=nc, strides=[w * c.itemsize for w in wc])
            assert np.array_equal(got, want), (na, pia, q, wa, wc, where)
            touched = np.zeros(span_c, bool)
            np.lib.stride_tricks.as_strided(touched, shape=nc, strides=list(wc))[...] = True
            assert np.all(c[~touched] == 55), "padding of C was written"
        # the oracle's slice nest (the reference's algorithm) on the same buffers; it accumulates into a zeroed C
        c_or = np.zeros(span_c, dtype)
        u = lambda t: np.asarray(t, np.uint64)
        assert oracle.run_raw(code, 0, q, p, a, u(na), u(wa), u(pia), b, u([len(b)]), c_or, u(nc), u(wc), u(pic)) == 0
        assert np.array_equal(np.lib.stride_tricks.as_strided(c_or, shape=nc, strides=[w * c_or.itemsize for w in wc]), want)
        # accumulate
        c = np.full(span_c, 3, dtype)
        ttv_b200.ttv_lowlevel(q, p, a, na, wa, pia, b, [len(b)], c, nc, wc, pic, flags=1)
        assert np.array_equal(np.lib.stride_tricks.as_strided(c, shape=nc, strides=[w * c.itemsize for w in wc]), want + 3)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128, np.int32, np.int64])
def test_padded_tensors_with_aligned_rows_take_vector_loads(dtype, monkeypatch):
    """general strides whose fastest free mode is contiguous in A and C and whose other strides are multiples of a
    16-byte vector (ttv_strided_vec_kernel: V consecutive outputs per thread): against the definition, with the
    vector form forced (TTV_B200_STRIDED_SCALAR=2; by default it is taken from ~2.4 M vector outputs on) and the
    thread-per-output form forced (=1) as a second opinion, odd n_q (partial last batch),
    accumulate, padding of C untouched, and a device buffer whose start is NOT vector-aligned (falls back to scalar)"""
    import torch
    rng = np.random.default_rng(33)
    cases = [
        # na, pia, q, wa, wc
        ((8, 5, 6), (1, 2, 3), 2, (1, 12, 64), (1, 12)),                    # rows padded 8 -> 12
        ((16, 37, 9, 3), (1, 2, 3, 4), 3, (1, 16, 600, 5600), (1, 20, 800)),  # q in the middle, C padded too
        ((64, 3, 21, 4), (1, 2, 3, 4), 3, (1, 64, 200, 4400), (1, 64, 192)),
        ((12, 10, 7), (1, 3, 2), 3, (1, 100, 12), (1, 16)),                  # layout (1,3,2): q = 3 sits in the middle
        ((4, 250, 8, 6), (1, 2, 3, 4), 2, (1, 4, 1024, 8192), (1, 4, 32)),   # a slice [:250] of a 256-extent mode
        ((6, 9, 5), (1, 2, 3), 2, (1, 8, 72), (1, 8)),                       # extent 6: only 8-byte vectors for 4-byte types
    ]
    for na, pia, q, wa, wc in cases:
        p = len(na)
        nc = ttv_b200.generate_output_shape(na, q); pic = ttv_b200.generate_output_layout(pia, q)
        x, a = _padded(rng, na, wa, dtype)
        b = rng.integers(-8, 9, na[q - 1]).astype(dtype)
        want = np.tensordot(x, b, axes=([q - 1], [0]))
        span_c = 1 + sum((n - 1) * w for n, w in zip(nc, wc))
        touched = np.zeros(span_c, bool)
        np.lib.stride_tricks.as_strided(touched, shape=nc, strides=list(wc))[...] = True
        results = []
        for scalar in ("2", "1"):                             # 2 = vector form wherever the strides allow, 1 = never
            monkeypatch.setenv("TTV_B200_STRIDED_SCALAR", scalar)
            c = np.full(span_c, 55, dtype)
            ttv_b200.ttv_lowlevel(q, p, a, na, wa, pia, b, [len(b)], c, nc, wc, pic)
            got = np.lib.stride_tricks.as_strided(c, shape=nc, strides=[w * c.itemsize for w in wc])
            assert np.array_equal(got, want), (na, pia, q, wa, wc, scalar)
            assert np.all(c[~touched] == 55), "padding of C was written"
            results.append(c)
            c = np.full(span_c, 3, dtype)
            ttv_b200.ttv_lowlevel(q, p, a, na, wa, pia, b, [len(b)], c, nc, wc, pic, flags=1)
            assert np.array_equal(np.lib.stride_tricks.as_strided(c, shape=nc, strides=[w * c.itemsize for w in wc]), want + 3)
        assert np.array_equal(results[0], results[1])
        monkeypatch.setenv("TTV_B200_STRIDED_SCALAR", "2")
        # device buffers one element past an aligned address: same strides, but no vector fits
        ta = torch.zeros(a.size + 1, dtype=torch.from_numpy(a).dtype, device="cuda")
        ta[1:] = torch.from_numpy(a).cuda()
        tb = torch.from_numpy(b).cuda()
        tc = torch.full((span_c + 1,), 55, dtype=ta.dtype, device="cuda")
        ttv_b200.ttv_lowlevel(q, p, ta[1:], na, wa, pia, tb, [len(b)], tc[1:], nc, wc, pic)
        c = tc[1:].cpu().numpy()
        assert np.array_equal(np.lib.stride_tricks.as_strided(c, shape=nc, strides=[w * c.itemsize for w in wc]), want)
        assert np.all(c[~touched] == 55)


@pytest.mark.parametrize("dtype", [np.float32, np.int32, np.float64, np.complex64])
def test_sliced_arrays_fuzz_both_general_stride_forms(dtype, monkeypatch):
    """seeded fuzz over slices of C-contiguous arrays (order 3..5, random cuts on every axis, the contiguous axis cut at
    multiples of the vector or anywhere), every q except the contiguous axis, host arrays and device tensors: the vector
    form forced (TTV_B200_STRIDED_SCALAR=2, falls back by itself where alignment forbids it), the thread-per-output form
    forced (=1) and the default rule must all give np.tensordot's result"""
    import torch
    rng = np.random.default_rng(20261017)
    for trial in range(24):
        p = int(rng.integers(3, 6))
        full = [int(rng.choice([3, 4, 6, 9, 12])) for _ in range(p - 1)] + [int(rng.choice([16, 36, 64, 132, 200]))]
        base = rng.integers(-8, 9, full).astype(dtype)
        if np.dtype(dtype).kind == "c":
            base = (base + 1j * rng.integers(-8, 9, full)).astype(dtype)
        cuts = []
        for ax, n in enumerate(full):
            if ax == p - 1:
                step = 4 if trial % 3 else 1                   # two trials out of three keep the rows vector-aligned
                lo = int(rng.integers(0, n // (2 * step))) * step
                hi = lo + max(step, int(rng.integers(1, (n - lo) // step + 1)) * step)
            else:
                lo = int(rng.integers(0, max(1, n // 2)))
                hi = int(rng.integers(lo + 1, n + 1))
            cuts.append(slice(lo, hi))
        x = base[tuple(cuts)]
        tx = torch.from_numpy(base).cuda()[tuple(cuts)]
        for q in range(1, p):
            b = rng.integers(-8, 9, x.shape[q - 1]).astype(dtype)
            want = np.tensordot(x, b, axes=([q - 1], [0]))
            for form in ("2", "1", "0"):
                monkeypatch.setenv("TTV_B200_STRIDED_SCALAR", form)
                got = ttv_b200.ttv(q, x, b)
                assert got.shape == want.shape and np.array_equal(got, want), (full, cuts, q, form, "host")
                tg = ttv_b200.ttv(q, tx, torch.from_numpy(b).cuda())
                assert np.array_equal(tg.cpu().numpy(), want), (full, cuts, q, form, "device")


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64])
def test_arrays_that_are_not_contiguous_are_read_in_place(dtype):
    """numpy / torch front end: transposes, slices and strided views for EVERY q -- also the cases
    1-7 in which the C-ABI ignores strides by default (TTV_B200_FLAG_HONOR_STRIDES) -- against np.tensordot"""
    import torch
    rng = np.random.default_rng(41)
    base = rng.integers(-8, 9, (9, 10, 11, 6)).astype(dtype)
    views = [base, base.transpose(2, 0, 3, 1), base[1:8, :, 2:9, :], base[:, ::2, :, ::3], base[:, 3, :, :], base[..., 0],
             np.asfortranarray(base)[:, 2:7], base[:, :, :, 1:2], base[::-1], base[2, :, :, 4].T]
    for x in views:
        for q in range(1, x.ndim + 1):
            b = rng.integers(-8, 9, x.shape[q - 1]).astype(dtype)
            want = np.tensordot(x, b, axes=([q - 1], [0]))
            got = ttv_b200.ttv(q, x, b)
            assert got.shape == want.shape and np.array_equal(got, want), (x.shape, x.strides, q)
            assert np.array_equal(ttv_b200.ttvpy.ttv(q, x, b), want)
            if min(x.strides) > 0:
                tx = torch.from_numpy(base).cuda().as_strided(x.shape, [s // x.itemsize for s in x.strides],
                                                              (x.__array_interface__["data"][0] - base.__array_interface__["data"][0]) // x.itemsize) \
                    if np.shares_memory(x, base) else torch.from_numpy(np.ascontiguousarray(x)).cuda()
                tg = ttv_b200.ttv(q, tx, torch.from_numpy(b).cuda())
                assert np.array_equal(tg.cpu().numpy(), want), (x.shape, x.strides, q, "device")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128, np.int32])
def test_padded_rows_when_q_is_the_contiguous_mode(dtype):
    """slices whose CONTIGUOUS axis is the contracted one (rows with a padded leading dimension): lane groups of
    ttv_strided_dot_kernel along the fibers -- vector widths 4 / 2 / 1 by the alignment of extent, strides and base
    address, group sizes 2..32, fibers longer than one pass of the group, fewer fibers than a warp holds, accumulate"""
    import torch
    rng = np.random.default_rng(43)
    bases = [rng.integers(-8, 9, shape).astype(dtype) for shape in [(50, 9, 300), (3, 2100), (7, 5, 3, 64), (1000, 36)]]
    if np.dtype(dtype).kind == "c":
        bases = [x + 1j * rng.integers(-8, 9, x.shape).astype(dtype) for x in bases]
    views = [bases[0][:, :, :250], bases[0][:, :, :248], bases[0][1:, 2:7, 3:36], bases[0][:, :, 1:10], bases[1][:, :2048],
             bases[1][:, 5:1030], bases[2][:, 1:4, :, :40], bases[3][:, :32], bases[3][::3, 4:36]]
    for x in views:
        q = x.ndim                                        # C-contiguous base: the last axis has stride 1
        b = rng.integers(-8, 9, x.shape[q - 1]).astype(dtype)
        want = np.tensordot(x, b, axes=([q - 1], [0]))
        got = ttv_b200.ttv(q, x, b)
        assert got.shape == want.shape and np.array_equal(got, want), (x.shape, x.strides)
        base = next(y for y in bases if np.shares_memory(x, y))
        off = (x.__array_interface__["data"][0] - base.__array_interface__["data"][0]) // x.itemsize
        tx = torch.from_numpy(base).cuda().as_strided(x.shape, [st // x.itemsize for st in x.strides], off)
        tg = ttv_b200.ttv(q, tx, torch.from_numpy(b).cuda())
        assert np.array_equal(tg.cpu().numpy(), want), (x.shape, x.strides, "device")
        # the C-like interface with the strides taken at their word (TTV_B200_FLAG_HONOR_STRIDES = 8), accumulate (1)
        p = x.ndim
        na = list(x.shape); wa = [st // x.itemsize for st in x.strides]
        pia = [int(m) + 1 for m in np.argsort(wa, kind="stable")]
        nc = ttv_b200.generate_output_shape(na, q); pic = ttv_b200.generate_output_layout(pia, q)
        wc = ttv_b200.generate_strides(nc, pic)
        flat = base.reshape(-1)[off:]
        c = np.full(want.size, 3, dtype)
        ttv_b200.ttv_lowlevel(q, p, flat, na, wa, pia, b, [len(b)], c, nc, wc, pic, flags=8 | 1)
        got = np.lib.stride_tricks.as_strided(c, shape=nc, strides=[w * c.itemsize for w in wc])
        assert np.array_equal(got, want + 3), (x.shape, x.strides, "lowlevel")


def test_async_device_calls_are_capturable_in_a_cuda_graph(oracle):
    """TTV_B200_FLAG_ASYNC calls on device pointers only enqueue work on the caller's stream (no allocation, no
    synchronisation once the workspace has its size), so a launch-bound sequence of products -- every mode of a small
    tensor, a split-n_q product with its reduce pass, a ttvs-like chain -- can be captured ONCE into a CUDA graph and
    replayed on new data with a single launch"""
    import torch
    rng = np.random.default_rng(77)
    na, pia = (24, 18, 20, 6), (2, 1, 4, 3)
    p = len(na)
    n = int(np.prod(na))
    a0, _ = random_case(rng, na, 1, np.float32)
    ta = torch.from_numpy(a0).cuda()
    tbs, tcs, meta = [], [], []
    for q in range(1, p + 1):
        tbs.append(torch.from_numpy(rng.integers(-8, 9, na[q - 1]).astype(np.float32)).cuda())
        nc = ttv_b200.generate_output_shape(na, q); pic = ttv_b200.generate_output_layout(pia, q)
        tcs.append(torch.zeros(n // na[q - 1], dtype=torch.float32, device="cuda"))
        meta.append((nc, ttv_b200.generate_strides(nc, pic), pic))
    wa = ttv_b200.generate_strides(na, pia)
    # a second stage on the result of q = 1 (a chain like ttvs), and a forced n_q split (workspace + reduce launch)
    nc1, wc1, pic1 = meta[0]
    tb2 = torch.from_numpy(rng.integers(-8, 9, nc1[0]).astype(np.float32)).cuda()
    nc2 = ttv_b200.generate_output_shape(nc1, 1); pic2 = ttv_b200.generate_output_layout(pic1, 1)
    tc2 = torch.zeros(int(np.prod(nc2)), dtype=torch.float32, device="cuda")
    tc_split = torch.zeros_like(tcs[2])
    stream = torch.cuda.Stream()

    def enqueue():
        for q in range(1, p + 1):
            nc, wc, pic = meta[q - 1]
            ttv_b200.ttv_lowlevel(q, p, ta, na, wa, pia, tbs[q - 1], [na[q - 1]], tcs[q - 1], nc, wc, pic, flags=2, stream=stream)
        ttv_b200.ttv_lowlevel(1, p - 1, tcs[0], nc1, wc1, pic1, tb2, [nc1[0]], tc2, nc2, ttv_b200.generate_strides(nc2, pic2), pic2,
                              flags=2, stream=stream)
        nc, wc, pic = meta[2]
        ttv_b200.ttv_lowlevel(3, p, ta, na, wa, pia, tbs[2], [na[2]], tc_split, nc, wc, pic, flags=2, ksplit=4, stream=stream)

    with torch.cuda.stream(stream):
        enqueue()                                    # warm-up outside the capture: sizes the workspace
    stream.synchronize()
    graph = torch.cuda.CUDAGraph()
    before = ttv_b200.launch_count()
    with torch.cuda.graph(graph, stream=stream):
        enqueue()
    assert ttv_b200.launch_count() - before >= p + 3          # p products + chain stage + split product and its reduce pass
    for trial in range(2):                           # replay on NEW contents of the same buffers
        a1, _ = random_case(rng, na, 1, np.float32)
        ta.copy_(torch.from_numpy(a1))
        for c in tcs + [tc2, tc_split]:
            c.fill_(-1)
        torch.cuda.synchronize()
        graph.replay()
        torch.cuda.synchronize()
        for q in range(1, p + 1):
            want = oracle.ttv(q, a1, na, pia, tbs[q - 1].cpu().numpy())
            assert np.array_equal(tcs[q - 1].cpu().numpy(), want), (trial, q)
        want1 = oracle.ttv(1, a1, na, pia, tbs[0].cpu().numpy())
        assert np.array_equal(tc2.cpu().numpy(), oracle.ttv(1, want1, nc1, pic1, tb2.cpu().numpy())), trial
        assert np.array_equal(tc_split.cpu().numpy(), oracle.ttv(3, a1, na, pia, tbs[2].cpu().numpy())), trial


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.complex64])
def test_host_tensors_are_streamed_in_chunks(dtype, oracle, monkeypatch):
    """host-pointer calls on large tensors stream A across PCIe in chunks of slabs of the slowest mode, kernels of one
    chunk overlapping the copy of the next: free split and n_q split, ragged last chunk, accumulate,
    several products per pass (ttv_multi); chunk size forced down to 1 MiB so that a few MB exercise the ring"""
    monkeypatch.setenv("TTV_B200_H2D_CHUNK_MB", "1")
    rng = np.random.default_rng(51)
    item = np.dtype(dtype).itemsize
    for na, pia in [((64, 50, 203), (1, 2, 3)), ((40, 30, 100, 9), (2, 1, 3, 4)), ((3000, 211), (1, 2)), ((211, 3000), (2, 1)),
                    ((16, 1, 300, 151, 1), (1, 2, 3, 4, 5))]:
        n = int(np.prod(na))
        assert n * item >= 2 << 20
        a = rng.integers(-8, 9, n).astype(dtype)
        bs, wants = [], []
        for q in range(1, len(na) + 1):
            b = rng.integers(-8, 9, na[q - 1]).astype(dtype)
            want = oracle.ttv(q, a, na, pia, b)
            bs.append(b); wants.append(want)
            before = ttv_b200.launch_count()
            got = run_lowlevel(q, a, na, pia, b)
            assert np.array_equal(got, want), (na, pia, q)
            if na[q - 1] > 1:         # (contracting an extent-1 slowest mode leaves nothing to split: plain path)
                assert ttv_b200.launch_count() - before >= 2, ("expected one launch per chunk", na, pia, q)
            c0 = np.full(want.size, 3, dtype)
            assert np.array_equal(run_lowlevel(q, a, na, pia, b, c0=c0, flags=1), want + 3), (na, pia, q, "accumulate")
        cs = ttv_b200.ttv_multi(list(range(1, len(na) + 1)), a, list(na), list(pia), bs)
        for q, (c, want) in enumerate(zip(cs, wants), 1):
            assert np.array_equal(c, want), (na, pia, q, "multi")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128, np.int32])
def test_fused_scatter_exchange_emulated_on_one_gpu(dtype, oracle):
    """ttv_b200_view_scatter + ttv_b200_reduce_slots (the n_q-split product fused with its exchange): `world` ranks are
    emulated on one device -- each "rank" owns a range of the contraction rows and writes its partial block by block into
    every owner's workspace; the owners sum their slots.  Peer memory over NVLink only changes what the pointers point to
    (that part runs in tools/sharded_bench.py --fused on N GPUs)."""
    import torch
    from ttv_b200.sharded import PeerExchange, split_range
    rng = np.random.default_rng(61)
    for (outer, nq, inner), world in [((1, 37, 5000), 3), ((1, 64, 1031), 4), ((5, 29, 700), 2), ((1, 8, 77), 8), ((3, 40, 256), 5)]:
        na, pia = (inner, nq, outer), (1, 2, 3)                  # A[outer][nq][inner] is this first-order tensor, q = 2
        a = rng.integers(-8, 9, outer * nq * inner).astype(dtype)
        b = rng.integers(-8, 9, nq).astype(dtype)
        want = oracle.ttv(2, a, na, pia, b)
        n = outer * inner
        blk = PeerExchange.block(n, world)
        assert blk % 256 == 0 and blk * world >= n
        a3 = torch.from_numpy(a).cuda().view(outer, nq, inner)
        tb = torch.from_numpy(b).cuda()
        ws = [torch.full((world * blk,), 99, dtype=a3.dtype, device="cuda") for _ in range(world)]
        before = ttv_b200.launch_count()
        for r in range(world):
            k0, kc = split_range(nq, world, r)
            a_r = a3[:, k0:k0 + kc, :].contiguous()
            ttv_b200.ttv_view_scatter(outer, kc, inner, a_r, tb[k0:k0 + kc].contiguous(), [w.data_ptr() for w in ws], r, blk)
        got = torch.empty(n, dtype=a3.dtype, device="cuda")
        for j in range(world):
            first = min(n, j * blk); cnt = max(0, min(blk, n - first))
            if cnt:
                ttv_b200.reduce_slots(ws[j], got[first:first + cnt], cnt, blk, world)
        torch.cuda.synchronize()
        assert ttv_b200.launch_count() - before >= world + 1
        assert np.array_equal(got.cpu().numpy(), want), ((outer, nq, inner), world, dtype)


def test_calls_from_several_host_threads(oracle, monkeypatch):
    """the C-ABI is callable from several host threads at once (ctypes drops the GIL): host-pointer calls share the staging
    buffers of their device and are serialised inside the library, device-pointer calls on their own streams share only
    the per-(device, stream) split-n_q workspace table.  Every result must still be the oracle's."""
    import threading
    import torch
    monkeypatch.setenv("TTV_B200_H2D_CHUNK_MB", "1")          # the larger cases take the chunk ring
    rng = np.random.default_rng(71)
    cases = []
    for na, pia, dtype in [((64, 50, 203), (1, 2, 3), np.float32), ((33, 47, 21), (3, 1, 2), np.float64), ((3000, 211), (1, 2), np.int32),
                           ((17, 9, 40, 23), (2, 1, 4, 3), np.complex64), ((211, 3000), (2, 1), np.int64), ((48, 40, 36), (1, 2, 3), np.float64)]:
        for q in range(1, len(na) + 1):
            a, b = random_case(rng, na, q, dtype)
            cases.append((q, a, na, pia, b, oracle.ttv(q, a, na, pia, b)))
    errors = []

    def host_worker(tid):
        try:
            for rep in range(3):
                for i in range(tid, len(cases), 2):
                    q, a, na, pia, b, want = cases[i]
                    if not np.array_equal(run_lowlevel(q, a, na, pia, b), want):
                        errors.append(("host", tid, na, pia, q))
        except Exception as e:                                # noqa: BLE001 - reported below
            errors.append(("host", tid, repr(e)))

    def device_worker(tid):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for rep in range(3):
                    for i in range(tid, len(cases), 2):
                        q, a, na, pia, b, want = cases[i]
                        ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
                        tc = torch.empty(want.size, dtype=ta.dtype, device="cuda")
                        nc = ttv_b200.generate_output_shape(na, q); pic = ttv_b200.generate_output_layout(pia, q)
                        ttv_b200.ttv_lowlevel(q, len(na), ta, na, ttv_b200.generate_strides(na, pia), pia, tb, [len(b)], tc, nc,
                                              ttv_b200.generate_strides(nc, pic), pic, ksplit=2 + (rep + tid) % 3,
                                              stream=stream.cuda_stream)
                        stream.synchronize()
                        if not np.array_equal(tc.cpu().numpy(), want):
                            errors.append(("device", tid, na, pia, q))
        except Exception as e:                                # noqa: BLE001
            errors.append(("device", tid, repr(e)))

    threads = [threading.Thread(target=host_worker, args=(t,)) for t in range(2)] + \
              [threading.Thread(target=device_worker, args=(t,)) for t in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]


def test_split_nq_from_several_threads_on_the_default_stream(oracle):
    """torch's default stream is shared by every host thread of a process.  The split-n_q variant is two launches (tiles ->
    partials, reduce -> C): the partials are a stream-ordered allocation of the call (api.cu: cudaMallocFromPoolAsync /
    cudaFreeAsync), so calls whose launches interleave on ONE stream -- A.tile, B.tile, A.reduce, B.reduce -- must still each
    read their own partials.  Asynchronous calls from four threads on the NULL stream, different shapes and partition counts."""
    import threading
    import torch
    rng = np.random.default_rng(72)
    cases = []
    for na, pia, dtype in [((64, 50, 203), (1, 2, 3), np.float32), ((3000, 211), (1, 2), np.int32), ((211, 3000), (2, 1), np.int64),
                           ((33, 47, 21), (3, 1, 2), np.float64), ((17, 9, 40, 23), (2, 1, 4, 3), np.complex64)]:
        for q in range(1, len(na) + 1):
            a, b = random_case(rng, na, q, dtype)
            ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
            cases.append((q, na, pia, ta, tb, oracle.ttv(q, a, na, pia, b)))
    torch.cuda.synchronize()
    errors = []
    barrier = threading.Barrier(4)

    def worker(tid):
        try:
            barrier.wait()
            for rep in range(6):
                outs = []
                for i in range(tid % 2, len(cases), 2):
                    q, na, pia, ta, tb, want = cases[i]
                    tc = torch.full((want.size,), 77, dtype=ta.dtype, device="cuda")
                    nc = ttv_b200.generate_output_shape(na, q); pic = ttv_b200.generate_output_layout(pia, q)
                    ttv_b200.ttv_lowlevel(q, len(na), ta, na, ttv_b200.generate_strides(na, pia), pia, tb, [int(tb.numel())], tc, nc,
                                          ttv_b200.generate_strides(nc, pic), pic, ksplit=2 + (rep + tid + i) % 5, flags=2, stream=0)
                    outs.append((i, tc))
                torch.cuda.synchronize()
                for i, tc in outs:
                    if not np.array_equal(tc.cpu().numpy(), cases[i][5]):
                        errors.append((tid, rep, cases[i][1], cases[i][2], cases[i][0]))
        except Exception as e:                                # noqa: BLE001 - reported below
            errors.append((tid, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_colt_kernel_tma_tensor_tiles(dtype, oracle, monkeypatch):
    """kernel="colt": A through shared memory by cp.async.bulk.tensor tiles (colt_kernel.cuh).  Rows narrower and wider than a
    box, rows that are not a power of two of words (zero-filled columns), contractions shorter / longer than a box and not a
    multiple of it (zero-filled rows), several slabs, n_q split across work items, few and many stages, accumulate."""
    rng = np.random.default_rng(23)
    name = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64", np.dtype(np.complex64): "c64",
            np.dtype(np.complex128): "c128", np.dtype(np.int32): "i32", np.dtype(np.int64): "i64"}[np.dtype(dtype)]
    vec = 16 // np.dtype(dtype).itemsize
    cases = [((64 * vec, 40, 3), (1, 2, 3), 2), ((1024 * vec, 33), (1, 2), 2), ((4 * vec, 300, 5), (1, 2, 3), 2), ((100 * vec, 7, 9), (1, 2, 3), 2),
             ((36 * vec, 5, 130, 4), (1, 2, 3, 4), 3), ((48, 17 * vec, 50), (2, 1, 3), 3), ((333 * vec, 257), (1, 2), 2), ((16 * vec, 2, 11), (1, 2, 3), 2)]
    for stages, stage_kb in (("4", "32"), ("2", "4"), ("8", "8")):
        monkeypatch.setenv("TTV_B200_COLT_STAGES", stages)
        monkeypatch.setenv("TTV_B200_COLT_STAGE_KB", stage_kb)
        for na, pia, q in cases:
            a, b = random_case(rng, na, q, dtype)
            want = oracle.ttv(q, a, na, pia, b)
            assert ttv_b200.plan(q, na, pia, dtype=name, kernel="colt")["kernel"] == 7
            for ks in (0, 3):
                assert np.array_equal(run_lowlevel(q, a, na, pia, b, kernel="colt", ksplit=ks), want), (na, pia, q, dtype, stages, ks)
            c0 = np.full(want.size, 3, dtype)
            assert np.array_equal(run_lowlevel(q, a, na, pia, b, c0=c0, kernel="colt", flags=1), want + 3)
    with pytest.raises(ttv_b200.TTVError):                          # rows that are not whole 16-byte vectors
        ttv_b200.plan(2, (64 * vec + 1, 9), (1, 2), dtype=name, kernel="colt") if vec > 1 else ttv_b200.plan(1, (9, 64), (1, 2), dtype=name, kernel="colt")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.int32, np.int64])
def test_dotp_kernel_fibers_of_two_elements(dtype, oracle, monkeypatch):
    """kernel="dotp": n_q = 2 with q the contiguous mode (dotp_kernel.cuh).  Even and odd numbers of fibers (the half vector
    at the end), fewer fibers than one vector / one tile / several tiles, both batch depths, accumulate; 16-byte elements
    are refused."""
    rng = np.random.default_rng(31)
    dt = np.dtype(dtype)
    name = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64", np.dtype(np.complex64): "c64", np.dtype(np.int32): "i32", np.dtype(np.int64): "i64"}[dt]
    cases = [((2, 1), (1, 2), 1), ((2, 2), (1, 2), 1), ((2, 3), (1, 2), 1), ((2, 1001), (1, 2), 1), ((2, 4096), (1, 2), 1), ((2, 2048 * 3 + 5), (1, 2), 1),
             ((7, 2, 11), (2, 1, 3), 2), ((5, 3, 2, 9), (3, 2, 1, 4), 3), ((2, 70001), (1, 2), 1), ((33, 2), (2, 1), 2)]
    for ku in ("8", "4"):
        monkeypatch.setenv("TTV_B200_DOTP_KU", ku)
        for na, pia, q in cases:
            a, b = random_case(rng, na, q, dtype)
            want = oracle.ttv(q, a, na, pia, b)
            assert ttv_b200.plan(q, na, pia, dtype=name, kernel="dotp")["kernel"] == 9
            assert ttv_b200.plan(q, na, pia, dtype=name)["kernel"] == 9
            assert np.array_equal(run_lowlevel(q, a, na, pia, b, kernel="dotp"), want), (na, pia, q, dtype, ku)
            assert np.array_equal(run_lowlevel(q, a, na, pia, b), want), (na, pia, q, dtype, ku)
            c0 = np.full(want.size, 3, dtype)
            assert np.array_equal(run_lowlevel(q, a, na, pia, b, c0=c0, kernel="dotp", flags=1), want + 3)
    with pytest.raises(ttv_b200.TTVError):
        ttv_b200.plan(1, (2, 64), (1, 2), dtype="c128", kernel="dotp")
    with pytest.raises(ttv_b200.TTVError):                          # three elements per fiber
        ttv_b200.plan(1, (3, 64), (1, 2), dtype=name, kernel="dotp")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.int32, np.int64])
def test_colf_kernel_rows_that_are_not_whole_vectors(dtype, oracle, monkeypatch):
    """kernel="colf": rows narrower than / not a multiple of a 16-byte vector, streamed flat as super-rows of V / gcd(inner, V)
    rows, a warp per slab or slab partition (colf_kernel.cuh).  Every gcd class (super-rows of 2 and 4 rows), one to 31
    vectors per super-row, contractions shorter than a super-row / a batch / several batches, rows past the last whole
    super-row (single slab only), several slabs, n_q split across warps (chosen and forced), short slabs side by side in one
    warp (with a last, partly filled group of slabs), accumulate; taken on its own from 256 bytes per slab."""
    rng = np.random.default_rng(37)
    dt = np.dtype(dtype)
    name = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64", np.dtype(np.complex64): "c64", np.dtype(np.int32): "i32", np.dtype(np.int64): "i64"}[dt]
    vec = 16 // dt.itemsize
    inners = [3, 5, 7, 9, 15, 31] + ([2, 6, 10, 14, 62] if vec == 4 else [])
    cases = []
    for i, inner in enumerate(inners):
        rows = vec // np.gcd(inner, vec)
        cases += [((inner, rows * (1000 + 7 * i), 3), (1, 2, 3), 2), ((inner, rows * 3), (1, 2), 2), ((inner, 1, rows * 2501), (1, 3, 2), 3),
                  ((inner, 4001 + i), (1, 2), 2), ((inner, rows - 1), (1, 2), 2), ((2, rows * 700, inner), (3, 2, 1), 2)]
    for na, pia, q in cases:
        a, b = random_case(rng, na, q, dtype)
        want = oracle.ttv(q, a, na, pia, b)
        assert ttv_b200.plan(q, na, pia, dtype=name, kernel="colf")["kernel"] == 10
        # rows of two 4-byte elements have a form of their own (ttv_colf2_kernel); the general kernel takes them as well
        for pair in (("1", "0") if (vec == 4 and 2 in na[:1] + na[-1:]) else ("1",)):
            monkeypatch.setenv("TTV_B200_COLF_PAIR", pair)
            for ks in (0, 1, 3):
                assert np.array_equal(run_lowlevel(q, a, na, pia, b, kernel="colf", ksplit=ks), want), (na, pia, q, dtype, ks, pair)
            c0 = np.full(want.size, 3, dtype)
            assert np.array_equal(run_lowlevel(q, a, na, pia, b, c0=c0, kernel="colf", flags=1), want + 3)
        monkeypatch.delenv("TTV_B200_COLF_PAIR")
    # slabs of at most one batch have a form of their own (ttv_colfs_kernel: b in registers); the general kernels take them as well
    for na in ((3, 4 * 6000, 2), (3, 4 * 30, 701), (2 if vec == 4 else 3, 256, 3000), (5, 64, 1001), (3, 32, 777), (2 if vec == 4 else 7, 32, 5003),
               (2 if vec == 4 else 3, 16, 40001), (3, 4 * 77, 33)):
        assert ttv_b200.plan(2, na, (1, 2, 3), dtype=name)["kernel"] == 10
        a, b = random_case(rng, na, 2, dtype)
        want = oracle.ttv(2, a, na, (1, 2, 3), b)
        for short in ("1", "0"):
            monkeypatch.setenv("TTV_B200_COLF_SHORT", short)
            assert np.array_equal(run_lowlevel(2, a, na, (1, 2, 3), b), want), (na, short)
            c0 = np.full(want.size, 5, dtype)
            assert np.array_equal(run_lowlevel(2, a, na, (1, 2, 3), b, c0=c0, flags=1), want + 5), (na, short)
        monkeypatch.delenv("TTV_B200_COLF_SHORT")
    with pytest.raises(ttv_b200.TTVError):                          # rows of whole vectors
        ttv_b200.plan(2, (2 * vec, 5000), (1, 2), dtype=name, kernel="colf")
    with pytest.raises(ttv_b200.TTVError):                          # more than 32 vectors per super-row
        ttv_b200.plan(2, (33, 5000), (1, 2), dtype=name, kernel="colf")
    with pytest.raises(ttv_b200.TTVError):                          # several slabs that do not start on a vector boundary
        ttv_b200.plan(2, (3, 4001, 2), (1, 2, 3), dtype=name, kernel="colf")


@pytest.mark.parametrize("dtype", [np.float32, np.int32])
def test_colf_tiny_slabs_of_two_element_rows(dtype, oracle, monkeypatch):
    """n_q = 2 .. 32 rows of two 4-byte elements: a slab is 1 .. 16 vectors, the lanes of a warp read consecutive vectors and the
    partial sums of a slab meet in a transposing butterfly (ttv_colf_tiny_kernel).  Every slab size, slab counts below / at / above
    one item of a warp (256 vectors) and one round of the grid, accumulate; the other kernels take the same shapes with
    TTV_B200_COLF_TINY=0."""
    rng = np.random.default_rng(41)
    name = "f32" if np.dtype(dtype) == np.float32 else "i32"
    for nq in (2, 4, 8, 16, 32):
        per_item = 8 * (32 // (nq // 2))
        for outer in (1, 3, per_item - 1, per_item, per_item + 1, 5 * per_item + 7, 40 * per_item + 3):
            for na, pia in (((2, nq, outer), (1, 2, 3)), ((outer, nq, 2), (3, 2, 1))):
                a, b = random_case(rng, na, 2, dtype)
                want = oracle.ttv(2, a, na, pia, b)
                pl = ttv_b200.plan(2, na, pia, dtype=name)
                assert (pl["kernel"], pl["ty"], pl["nu"]) == (10, nq // 2, 32 // (nq // 2)), (na, pl)
                assert np.array_equal(run_lowlevel(2, a, na, pia, b), want), (na, pia, dtype)
                c0 = np.full(want.size, 3, dtype)
                assert np.array_equal(run_lowlevel(2, a, na, pia, b, c0=c0, flags=1), want + 3), (na, pia, dtype)
    na = (2, 16, 70001)                                             # several rounds of the persistent grid
    a, b = random_case(rng, na, 2, dtype)
    want = oracle.ttv(2, a, na, (1, 2, 3), b)
    assert np.array_equal(run_lowlevel(2, a, na, (1, 2, 3), b), want)
    monkeypatch.setenv("TTV_B200_COLF_TINY", "0")
    assert ttv_b200.plan(2, na, (1, 2, 3), dtype=name)["ty"] != 8
    assert np.array_equal(run_lowlevel(2, a, na, (1, 2, 3), b), want)


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_streamk_kernel_tiny_inner_long_contraction(dtype, oracle, monkeypatch):
    """kernel="streamk": rows of a few elements under a long contraction, staged through shared memory by bulk copies with the
    threads along n_q (streamk_kernel.cuh).  Rows of 2 .. 64 bytes, contractions that are not a multiple of a stage, tensors
    and vectors whose last bytes no 16-byte bulk copy covers (plain-load tail), several slabs, n_q split across CTAs (forced and
    chosen), one and many stages per partition, accumulate."""
    rng = np.random.default_rng(29)
    dt = np.dtype(dtype)
    name = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64", np.dtype(np.complex64): "c64",
            np.dtype(np.complex128): "c128", np.dtype(np.int32): "i32", np.dtype(np.int64): "i64"}[dt]
    imax = 64 // dt.itemsize
    inners = sorted({2, 3, min(5, imax), min(7, imax), imax})
    cases = []
    for i, inner in enumerate(inners):
        cases += [((inner, 4099 + 2 * i, 3), (1, 2, 3), 2), ((inner, 4096, 1), (1, 2, 3), 2), ((2, 9001, inner), (3, 2, 1), 2), ((inner, 1, 20011), (1, 3, 2), 3)]
    for stage_kb in ("32", "1", "6"):
        monkeypatch.setenv("TTV_B200_STREAMK_STAGE_KB", stage_kb)
        for na, pia, q in cases:
            a, b = random_case(rng, na, q, dtype)
            want = oracle.ttv(q, a, na, pia, b)
            assert ttv_b200.plan(q, na, pia, dtype=name, kernel="streamk")["kernel"] == 8
            for ks in (0, 1, 3):
                assert np.array_equal(run_lowlevel(q, a, na, pia, b, kernel="streamk", ksplit=ks), want), (na, pia, q, dtype, stage_kb, ks)
            c0 = np.full(want.size, 3, dtype)
            assert np.array_equal(run_lowlevel(q, a, na, pia, b, c0=c0, kernel="streamk", flags=1), want + 3)
    with pytest.raises(ttv_b200.TTVError):                          # rows wider than 64 bytes
        ttv_b200.plan(2, (imax + 1, 5000), (1, 2), dtype=name, kernel="streamk")
    with pytest.raises(ttv_b200.TTVError):                          # a short contraction
        ttv_b200.plan(2, (3, 100), (1, 2), dtype=name, kernel="streamk")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.complex64, np.complex128])
def test_single_kernel_exchange_emulated_on_one_gpu(dtype, oracle, monkeypatch):
    """ttv_b200_view_exchange: product + scatter + in-kernel flag barrier + sum of the slots in ONE kernel per rank.  The
    `world` ranks are `world` kernels on separate streams of one device; only the last few CTAs of each kernel to finish stay
    for the barrier (max_ctas caps them, so that the waiting CTAs of all emulated ranks fit on the device together with the
    CTAs still computing).  Several rounds: tokens grow, the two workspace halves alternate, the arrival counter resets
    itself.  The timeout is short so that a protocol bug fails instead of hanging the device."""
    import torch
    from ttv_b200.sharded import PeerExchange, split_range
    monkeypatch.setenv("TTV_B200_EXCHANGE_TIMEOUT_MS", "3000")
    rng = np.random.default_rng(62)
    token = 0
    for (outer, nq, inner), world in [((1, 37, 5000), 3), ((1, 64, 1031), 4), ((5, 29, 700), 2), ((1, 8, 77), 8), ((3, 40, 256), 5), ((1, 200, 40000), 2)]:
        na, pia = (inner, nq, outer), (1, 2, 3)
        n = outer * inner
        blk = PeerExchange.block(n, world)
        tdt = torch.from_numpy(np.zeros(1, dtype)).dtype
        ws = [torch.full((2 * world * blk,), 99, dtype=tdt, device="cuda") for _ in range(world)]        # two halves per rank
        flags = [torch.zeros(16, dtype=torch.int32, device="cuda") for _ in range(world)]
        scratch = [torch.zeros(4, dtype=torch.int32, device="cuda") for _ in range(world)]
        streams = [torch.cuda.Stream() for _ in range(world)]
        torch.cuda.synchronize()
        for rnd in range(3):
            a = rng.integers(-8, 9, outer * nq * inner).astype(dtype)
            b = rng.integers(-8, 9, nq).astype(dtype)
            want = oracle.ttv(2, a, na, pia, b)
            a3 = torch.from_numpy(a).cuda().view(outer, nq, inner)
            tb = torch.from_numpy(b).cuda()
            parts = []
            for r in range(world):
                k0, kc = split_range(nq, world, r)
                parts.append((a3[:, k0:k0 + kc, :].contiguous(), tb[k0:k0 + kc].contiguous()))
            got = torch.full((n,), 55, dtype=tdt, device="cuda")
            token += 1
            half = (token % 2) * world * blk
            itemsize = np.dtype(dtype).itemsize
            torch.cuda.synchronize()
            for r in range(world):
                first = min(n, r * blk); cnt = max(0, min(blk, n - first))
                with torch.cuda.stream(streams[r]):
                    ttv_b200.ttv_view_exchange(outer, parts[r][0].shape[1], inner, parts[r][0], parts[r][1],
                                               [w.data_ptr() + half * itemsize for w in ws], [f.data_ptr() for f in flags], r, blk,
                                               got[first:first + cnt] if cnt else None, token, scratch[r], max_ctas=max(1, 96 // world),
                                               stream=streams[r].cuda_stream)
            torch.cuda.synchronize()
            assert all(int(s[2].item()) == 0 for s in scratch), "an emulated rank timed out waiting for the others"
            assert np.array_equal(got.cpu().numpy(), want), ((outer, nq, inner), world, dtype, rnd)


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_split_nq_with_many_partitions_and_few_outputs(dtype, oracle):
    """the second pass of the split-n_q variant has two forms: a thread per output (many outputs) and a CTA per output with a
    shared-memory tree (ttv_reduce_wide_kernel: 64+ partitions, <= 8192 outputs -- what [1, 2^26, 4] of the asymmetric family
    takes at full size).  Forced partition counts on both sides of the switch, accumulate, odd output counts."""
    rng = np.random.default_rng(31)
    for na, pia, q in [((4, 50000), (1, 2), 2), ((3, 40000, 5), (1, 2, 3), 2), ((7, 30011), (1, 2), 2), ((30011, 9), (2, 1), 1), ((1, 70000, 2), (1, 2, 3), 2)]:
        a, b = random_case(rng, na, q, dtype)
        want = oracle.ttv(q, a, na, pia, b)
        for ks in (63, 64, 100, 777):
            assert np.array_equal(run_lowlevel(q, a, na, pia, b, ksplit=ks), want), (na, pia, q, dtype, ks)
        c0 = np.full(want.size, 3, dtype)
        assert np.array_equal(run_lowlevel(q, a, na, pia, b, c0=c0, ksplit=200, flags=1), want + 3), (na, pia, q, dtype, "accumulate")
