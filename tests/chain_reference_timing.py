#!/usr/bin/env python
"""tests/chain_reference_timing.py -- times the REFERENCE's own compiled Python module (oracle/_ref/ttvpy_ref*.so, built by
oracle/Makefile from /root/reference/ttvpy/src/wrapped_ttv.cpp, OpenMP, no BLAS) on the ttvs chain, on the host cores.
Checker-side measurement (lives under tests/ because only tests/, smoke() and bench.py's cpu_baseline leg may execute
anything under oracle/); the GPU side of the same table is tools/chain_bench.py.

    python tests/chain_reference_timing.py [--shape 256,256,256,128]
"""
from __future__ import annotations

import argparse
import glob
import importlib.util
import json
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_reference():
    cands = glob.glob(os.path.join(ROOT, "oracle", "_ref", "ttvpy_ref*.so"))
    if not cands:
        raise SystemExit("oracle/_ref/ttvpy_ref*.so is missing: run `make -C oracle` where /root/reference exists")
    spec = importlib.util.spec_from_file_location("ttvpy_ref", cands[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="256,256,256,128")
    args = ap.parse_args()
    shape = [int(x) for x in args.shape.split(",")]
    p = len(shape)
    ref = load_reference()
    rng = np.random.default_rng(0)
    A = rng.uniform(-1, 1, shape)
    for q in (1, p):
        vecs = [np.linspace(-1, 1, shape[r]) for r in range(p) if r != q - 1]
        for order in ("optimal", "backward", "forward"):
            ref.ttvs(q, A, vecs, order)                                   # warm-up (thread pool, page faults)
            t0 = time.perf_counter(); ref.ttvs(q, A, vecs, order); ms = (time.perf_counter() - t0) * 1e3
            print(json.dumps({"shape": shape, "q": q, "order": order, "ms_ref": ms, "cores": os.cpu_count()}), flush=True)


if __name__ == "__main__":
    main()
