#!/usr/bin/env python
"""tools/chain_bench.py -- ttvpy.ttvs: the chain of p-1 products that leaves mode q, fp64 like the
reference's binding (ttvpy/src/wrapped_ttv.cpp:83-198).

  device   A and the vectors resident in HBM, intermediates stay there; CUDA events around the whole chain
  host     numpy in, numpy out: A crosses PCIe once, only the final vector comes back
  (the reference module's own timings on the host cores -- profiles/r01_ttvs_chain.txt, first block -- were taken once by
   tests/chain_reference_timing.py; nothing under oracle/ is touched from here)

GB/s = sum over the p-1 steps of the algorithmic bytes of that step (tensor + vector + result) / time.
    python tools/chain_bench.py [--shape 256,256,256,128] [--reps 5]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import ttv_b200  # noqa: E402
from ttv_b200 import ttvpy  # noqa: E402


def chain_bytes(q, shape, order, item=8):
    shape = list(shape)
    total = 0
    for mode, _ in ttvpy.chain_plan(q, shape, order):
        n = int(np.prod(shape, dtype=object))
        nq = shape[mode - 1]
        total += item * (n + nq + n // nq)
        del shape[mode - 1]
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="256,256,256,128")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-ref", action="store_true", help="accepted for old command lines; there is no reference leg any more")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "chain.jsonl"))
    args = ap.parse_args()
    shape = [int(x) for x in args.shape.split(",")]
    p = len(shape)
    n = int(np.prod(shape))
    dev = torch.empty(n, dtype=torch.float64, device="cuda")
    ttv_b200.fill(dev, 0x77170001)
    A_dev = dev.view(*shape)
    A_host = A_dev.cpu().numpy()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "a") as f:
        for q in (1, p):
            vec_host = [np.linspace(-1, 1, shape[r]).astype(np.float64) for r in range(p) if r != q - 1]
            vec_dev = [torch.from_numpy(v).cuda() for v in vec_host]
            for order in ("optimal", "backward", "forward"):
                byt = chain_bytes(q, shape, order)
                out = ttvpy.ttvs(q, A_dev, vec_dev, order)                       # warm-up
                torch.cuda.synchronize()
                ts = []
                for _ in range(args.reps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); out = ttvpy.ttvs(q, A_dev, vec_dev, order); e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms_dev = sorted(ts)[len(ts) // 2]
                t0 = time.perf_counter(); got = ttvpy.ttvs(q, A_host, vec_host, order); ms_host = (time.perf_counter() - t0) * 1e3
                rec = {"shape": shape, "q": q, "order": order, "chain_bytes": byt, "ms_device": ms_dev, "gbs_device": byt / ms_dev / 1e6,
                       "ms_host": ms_host, "gbs_host": byt / ms_host / 1e6}
                # the same chain captured into a CUDA graph (ttvpy.CapturedTtvs): one launch per replay
                plan = ttvpy.CapturedTtvs(q, A_dev, vec_dev, order)
                plan.replay(); torch.cuda.synchronize()
                ts = []
                for _ in range(max(args.reps, 10)):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); plan.replay(); e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms_graph = sorted(ts)[len(ts) // 2]
                rec.update(ms_graph=ms_graph, gbs_graph=byt / ms_graph / 1e6,
                           graph_matches=bool(torch.equal(plan.result, ttvpy.ttvs(q, A_dev, vec_dev, order))))
                del plan
                print(json.dumps(rec), flush=True)
                f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
