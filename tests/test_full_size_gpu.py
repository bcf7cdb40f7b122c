"""Parity at FULL size for the plans the chooser only picks on large tensors (ttv_b200/csrc/plan.cpp keys on sizes: the
big-slab STREAM form, n_q split + direct b, 128-thread CTAs, the (1,16) deep batch, COLX / COLW, DOTF on 16-byte elements,
the peeled DOT, fibers cut into pieces, ...).  Small shapes reach those kernel FAMILIES when forced, but not these exact
plans; here every one of them runs on its BASELINE-sized tensor and sampled outputs are compared with a host long-double dot
on regenerated fibers (ttv_b200/selfcheck.py; int32 bit-exact).  bench.py's sweep leg does the same for all 254 named
products; this is the subset that pins one product per chooser branch in the test suite.

The reference's own grid only reaches extents {2,4,8}^p (test/src/gtest_tlib_ttv.cpp:192-425)."""
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# (config name, dtype, q) -> what the plan must say, so that the test really exercises the branch it names
BRANCHES = [
    ("cfg1", "f32", 1, dict(kernel=1)),                                  # DOT, four vectors per lane (tensor < 2 GiB)
    ("cfg1", "f32", 2, dict(kernel=2, nu=1, ku=16)),                     # column GEMV, deep batch (1,16): few CTAs
    ("cfg1", "f32", 3, dict(kernel=2, nu=1, ku=16)),
    ("sym4", "f32", 2, dict(kernel=2, nu=1, ku=8)),                      # the bench kernel
    ("sym2", "f32", 1, dict(kernel=1, ty=32, ksplit_gt=1)),              # warp-per-fiber DOT, 256 KB fibers cut into pieces
    ("sym2", "f32", 2, dict(kernel=2, ksplit_gt=1)),                     # split n_q two-pass (64 tiles only)
    ("sym3", "f32", 1, dict(kernel=1)),                                  # peeled DOT: fibers of 1625 floats
    ("sym3", "f32", 2, dict(kernel=4)),                                  # COLX, CTA form: rows of 1625 floats
    ("sym3", "f32", 3, dict(kernel=4)),
    ("sym5", "f32", 1, dict(kernel=5)),                                  # DOTF: fibers of 84 floats
    ("sym7", "f32", 1, dict(kernel=3)),                                  # STREAM: fibers of 23 floats
    ("sym7", "f32", 2, dict(kernel=3)),                                  # STREAM: 23 x 23 slabs
    ("sym7", "f32", 3, dict(kernel=3, threads_gt=256)),                  # STREAM, big slab alone in its stage (23 x 529)
    ("sym7", "f32", 5, dict(kernel=4)),                                  # COLX warp form (COLW): 23 rows
    ("sym7", "f32", 7, dict(kernel=4)),
    ("cfg5/8", "f64", 1, dict(kernel=1)),                                # cfg5 slab, DOT on 16 KB fibers
    ("cfg5/8", "f64", 3, dict(kernel=2)),                                # cfg5 slab, one big column GEMV
    ("sym2d", "f64", 2, dict(kernel=2, ksplit_gt=1)),
    ("sym5d", "f64", 1, dict(kernel=3)),                                 # fibers of 73 doubles: STREAM from 48 outputs per stage
    ("sym5d", "f64", 3, dict(kernel=4)),                                 # 73^5: COLX warp form below 48 rows per phase lane
    ("sym7d", "f64", 3, dict(kernel=3)),                                 # 21 x 441 doubles: 74 KB slab
    ("asym4", "i32", 2, None),                                           # int32 bit-exact on the asymmetric family
    ("asym6", "i32", 3, dict(kernel=10, ksplit_gt=1)),                   # n_q = 2^20, rows of 6: COLF, super-rows of two rows, n_q split over warps
    ("asym7", "f32", 2, dict(kernel=10, ksplit_gt=1)),                   # rows of two floats under n_q = 2^17: COLF
    ("asym5n", "f32", 2, dict(kernel=10, ksplit=1)),                     # 8 388 608 slabs of 128 x 2 floats: COLF, four slabs per warp
    ("asym5n", "i32", 2, dict(kernel=10, ksplit=1)),
    ("asym3n", "f32", 1, dict(kernel=9)),                                # DOTP: fibers of two floats
    ("asym10", "f32", 2, dict(kernel=10, ty=1, nu=32)),                  # 805 306 368 slabs of 2 x 2 floats: COLF, tiny slabs (a vector per slab)
    ("asym10", "i32", 2, dict(kernel=10, ty=1, nu=32)),
    ("asym10", "i32", 1, dict(kernel=9)),                                # n_q = 2: a third of the traffic is writes
    ("asym10", "f32", 10, None),
    ("asym8", "f32", 2, None),
    ("cplx5", "c128", 5, dict(kernel=5)),                                # DOTF on 16-byte elements
    ("cplx6", "c128", 2, None),
    ("cx6L", "c128", 4, dict(threads=128)),                              # rows of 625 vectors: 128-thread CTAs
    ("cx4R1", "c64", 2, None),
]


def _find(name, dt, q):
    from ttv_b200.workloads import configs
    for cfg in configs("named"):
        if cfg[0] == name and cfg[1] == dt and cfg[4] == q:
            return cfg
    raise KeyError((name, dt, q))


@pytest.fixture(scope="module")
def arenas():
    from ttv_b200.measure import Arena
    return Arena(int(17.5e9)), Arena(int(8.8e9))


@pytest.mark.parametrize("name,dt,q,want", BRANCHES, ids=[f"{n}-{d}-q{q}" for n, d, q, _ in BRANCHES])
def test_full_size_plan_matches_host_dot_on_sampled_fibers(arenas, name, dt, q, want):
    import ttv_b200
    from ttv_b200.measure import measure_config
    _, _, na, pia, _ = _find(name, dt, q)[:5]
    pl = ttv_b200.plan(q, na, pia, dtype=dt)
    for key, val in (want or {}).items():
        if key.endswith("_gt"):
            assert pl[key[:-3]] > val, (key, pl)
        else:
            assert pl[key] == val, (key, pl)
    r = measure_config(dt, na, pia, q, reps=1, warmup=0, samples=48, arena_a=arenas[0], arena_c=arenas[1],
                       rng=np.random.default_rng(zlib.crc32(f"{name}-{dt}-{q}".encode())))
    assert r["checked"] >= 4 and r["failures"] == 0, (name, dt, q, pl, r)
