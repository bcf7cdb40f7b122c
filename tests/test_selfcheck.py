"""ttv_b200/selfcheck.py is the product's own sampled self-check (bench.py, tools/sweep.py, the full-size GPU tests): a numpy
restatement of the fill generator + a host long-double dot per sampled output.  Here it is pinned against the oracle on the
CPU: the generator element by element, the expected values / tolerances against oracle.ttv, and the small-tensor numpy
product (full_product) for every layout."""
import itertools

import numpy as np

from ttv_b200 import selfcheck as sc
from ttv_b200 import workloads


def test_generator_matches_the_oracle_fill(oracle):
    for dt in sc.NP_DTYPE:
        for first in (0, 12345, (1 << 33) + 7):
            want = oracle.fill(dt, 1000, 0x77170001, first=first)
            got = sc.synth(dt, 0x77170001, np.arange(first, first + 1000, dtype=np.uint64))
            assert got.dtype == want.dtype and np.array_equal(got, want), (dt, first)


def test_expected_values_match_the_oracle_product(oracle):
    for dt in sc.NP_DTYPE:
        for na, pia in [((6, 5, 7), (2, 3, 1)), ((4, 9, 3, 5), (1, 2, 3, 4)), ((11, 13), (2, 1))]:
            n = int(np.prod(na))
            a = oracle.fill(dt, n, 5)
            for q in range(1, len(na) + 1):
                b = oracle.fill(dt, na[q - 1], 9)
                want = oracle.ttv(q, a, na, pia, b)
                view = sc.view_of(na, pia, q)
                assert view[0] * view[1] * view[2] == n
                vals, tols = sc.expected(dt, view, np.arange(want.size), 5, b)
                if dt in ("i32", "i64"):
                    assert [int(v) for v in vals] == want.tolist() and not any(tols)
                else:
                    err = np.abs(np.asarray(vals) - want)
                    assert (err <= np.asarray(tols)).all(), (dt, na, pia, q)
                # a sharded rank regenerates fibers from GLOBAL indices: offsetting C by c_first must give the same values
                half = want.size // 2
                vals2, _ = sc.expected(dt, view, np.arange(want.size - half), 5, b, c_first=half)
                assert np.allclose(np.asarray(vals2), np.asarray(vals[half:]), rtol=0, atol=0)


def test_full_product_matches_the_oracle_for_every_layout(oracle):
    rng = np.random.default_rng(3)
    for na in [(5, 4, 6), (3, 7, 2, 4)]:
        for pia in itertools.permutations(range(1, len(na) + 1)):
            for dtype in (np.int32, np.float64, np.complex64):
                a = rng.integers(-6, 7, int(np.prod(na))).astype(dtype)
                for q in range(1, len(na) + 1):
                    b = rng.integers(-6, 7, na[q - 1]).astype(dtype)
                    assert np.array_equal(sc.full_product(a, na, pia, q, b), oracle.ttv(q, a, na, pia, b)), (na, pia, q)


def test_named_workloads_are_the_254_products_of_baseline_json():
    named = workloads.configs("named")
    assert len(named) == 254 and len(workloads.configs("all")) == 122 and len(workloads.configs("cplxall")) == 90 and len(workloads.configs("asym2")) == 42
    assert {len(c[2]) for c in named if c[0].startswith("asym")} == set(range(2, 11))      # BASELINE configs[2]: every order p = 2..10
    assert ("cfg1", "f32", [512, 512, 512], [1, 2, 3], 2) in named                      # BASELINE configs[0]
    assert sum(1 for c in named if c[1] == "i32") == 45 and {c[1] for c in named} == {"f32", "f64", "c64", "c128", "i32"}
    for name, dt, na, pia, q, *rest in named:
        assert sorted(pia) == list(range(1, len(na) + 1)) and 1 <= q <= len(na)
        assert workloads.algo_bytes(dt, na, q) == workloads.SIZE[dt] * (int(np.prod(na, dtype=object)) * (na[q - 1] + 1) // na[q - 1] + na[q - 1])
