"""Generates tests/golden/ttv_golden.npz from the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile from
/root/reference/include) and tests/golden/ttvpy_golden.npz from the reference's own Python module
(oracle/_ref/ttvpy_ref*.so, built from /root/reference/ttvpy/src/wrapped_ttv.cpp).

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The fixtures are small (a few hundred KB) and committed; the GPU box never needs /root/reference.

Each case stores na, pia, q, flat a, b and the reference's flat c.  `exact` marks integer-valued data (every policy of
the reference agrees bit for bit); the others hold values in [-1,1) and are compared with the stated tolerance.
The policy that produced c cycles through all 17 combinations and both builds (OpenMP-only, OpenBLAS).
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.oracle import Oracle, Reference, REF_COMBOS  # noqa: E402
from conftest import all_layouts, random_case, real_case  # noqa: E402


def main():
    oracle = Oracle()
    libs = [Reference()] + ([Reference(blas=True)] if Reference.available(blas=True) else [])
    rng = np.random.default_rng(20261017)
    out = {}
    n = 0
    dtypes = [np.float32, np.float64, np.complex64, np.complex128, np.int32, np.int64]
    shapes = {2: [(5, 7), (16, 3), (2, 33)],
              3: [(4, 3, 2), (3, 5, 7), (8, 2, 9), (2, 17, 4)],
              4: [(2, 3, 4, 5), (5, 2, 2, 6), (3, 3, 3, 3)],
              5: [(2, 3, 2, 3, 4), (4, 2, 3, 2, 2)],
              6: [(2, 2, 3, 2, 2, 3)]}
    for order, shape_list in shapes.items():
        layouts = all_layouts(order)
        if len(layouts) > 6:      # first-order, last-order and a seeded sample of the rest
            pick = [0, len(layouts) - 1] + sorted(rng.choice(len(layouts) - 2, 4, replace=False) + 1)
            layouts = [layouts[i] for i in pick]
        for na in shape_list:
            for pia in layouts:
                for q in range(1, order + 1):
                    dtype = dtypes[n % len(dtypes)]
                    exact = (n // len(dtypes)) % 2 == 0 or np.dtype(dtype).kind in "iu"
                    a, b = (random_case if exact else real_case)(rng, na, q, dtype)
                    lib = libs[n % len(libs)]
                    combo = REF_COMBOS[n % len(REF_COMBOS)]
                    c = lib.ttv(q, a, na, pia, b, combo=combo, helpers=oracle)
                    out[f"na_{n}"] = np.asarray(na, np.int64); out[f"pia_{n}"] = np.asarray(pia, np.int64)
                    out[f"q_{n}"] = np.int64(q); out[f"a_{n}"] = a; out[f"b_{n}"] = b; out[f"c_{n}"] = c
                    out[f"exact_{n}"] = np.bool_(exact)
                    out[f"combo_{n}"] = np.asarray("/".join(combo) + ("+openblas" if lib.blas else ""))
                    n += 1
    out["count"] = np.int64(n)
    np.savez_compressed(os.path.join(HERE, "ttv_golden.npz"), **out)
    print(f"wrote {n} TTV cases")

    # ---- the reference's python module: ttv and ttvs ---------------------------------------------------------------
    cand = [f for f in os.listdir(os.path.join(ROOT, "oracle", "_ref")) if f.startswith("ttvpy_ref")]
    if not cand:
        print("no ttvpy_ref module; skipping ttvpy fixtures")
        return
    spec = importlib.util.spec_from_file_location("ttvpy_ref", os.path.join(ROOT, "oracle", "_ref", cand[0]))
    ttvpy_ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ttvpy_ref)
    py = {}
    m = 0
    for shape in [(3, 2, 4), (3, 2, 4, 5), (5, 5, 5, 5), (2, 6), (4, 3, 2, 3, 2)]:
        p = len(shape)
        A = rng.integers(-4, 5, shape).astype(np.float64)
        for q in range(1, p + 1):
            b = rng.integers(-4, 5, shape[q - 1]).astype(np.float64)
            py[f"ttv_A_{m}"] = A; py[f"ttv_b_{m}"] = b; py[f"ttv_q_{m}"] = np.int64(q)
            py[f"ttv_C_{m}"] = np.ascontiguousarray(ttvpy_ref.ttv(q, A, b))
            bs = [rng.integers(-3, 4, shape[r]).astype(np.float64) for r in range(p) if r != q - 1]
            for j, bj in enumerate(bs):
                py[f"ttvs_b_{m}_{j}"] = bj
            for order in ("forward", "backward", "optimal"):
                py[f"ttvs_C_{m}_{order}"] = np.ascontiguousarray(ttvpy_ref.ttvs(q, A, bs, order))
            m += 1
    py["count"] = np.int64(m)
    np.savez_compressed(os.path.join(HERE, "ttvpy_golden.npz"), **py)
    print(f"wrote {m} ttvpy cases")


if __name__ == "__main__":
    main()
