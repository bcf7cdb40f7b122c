#!/usr/bin/env python
"""tools/probe/multi_device_probe.py -- one HOST tensor over 1, 2, ... GPUs of this process (ttv_b200_run_devices).

    python tools/probe/multi_device_probe.py [--gib 16] [--pageable]

The 256^3 x m fp32 tensor (m sized for --gib) sits in pinned host memory (ttv_b200_host_alloc) -- or pageable with --pageable --
and every mode q = 1..4 is contracted through the drop-in call with host pointers, once per device list.  Prints the
effective GB/s of the call (algorithmic bytes / wall time), i.e. the per-call e2e rate a host-resident caller sees, and
checks the results of every device list against the single-GPU result."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
import ttv_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=16.0)
    ap.add_argument("--pageable", action="store_true")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    n_dev = torch.cuda.device_count()
    m = max(8, int(args.gib * 2 ** 30 / 4 / 256 ** 3) // 8 * 8)
    na, pia = [256, 256, 256, m], [1, 2, 3, 4]
    n = int(np.prod(na))
    a = np.empty(n, np.float32) if args.pageable else ttv_b200.pinned_empty(n, np.float32)
    chunk = 1 << 24
    pattern = np.random.default_rng(1).uniform(-1, 1, chunk).astype(np.float32)
    for s0 in range(0, n, chunk):
        a[s0:s0 + chunk] = pattern[: min(chunk, n - s0)]
    print(f"tensor {na} fp32 = {n * 4 / 2 ** 30:.1f} GiB in {'pageable' if args.pageable else 'pinned'} host memory, {n_dev} GPU(s) visible", flush=True)
    lists = [[0]] + [list(range(g)) for g in (2, 4, 8) if g <= n_dev]
    base = {}
    for devices in lists:
        for q in (1, 2, 3, 4):
            b = np.random.default_rng(10 + q).uniform(-1, 1, na[q - 1]).astype(np.float32)
            nc = ttv_b200.generate_output_shape(na, q); pic = ttv_b200.generate_output_layout(pia, q)
            c = ttv_b200.pinned_empty(n // na[q - 1], np.float32)
            args_ = (q, 4, a, na, ttv_b200.generate_strides(na, pia), pia, b, [na[q - 1]], c, nc, ttv_b200.generate_strides(nc, pic), pic)
            ttv_b200.ttv_lowlevel_devices(devices, *args_)                      # warm-up: staging buffers, streams
            t0 = time.perf_counter()
            for _ in range(args.reps):
                ttv_b200.ttv_lowlevel_devices(devices, *args_)
            dt = (time.perf_counter() - t0) / args.reps
            byt = 4 * (n + na[q - 1] + n // na[q - 1])
            if len(devices) == 1:
                base[q] = c.copy()
                ok = "reference"
            else:
                tol = na[q - 1] * float(np.finfo(np.float32).eps) * float(np.abs(b).sum())
                ok = "ok" if float(np.abs(c - base[q]).max()) <= tol else "MISMATCH"
            print(f"devices={len(devices)} q={q}: {dt * 1e3:8.1f} ms  {byt / dt / 1e9:7.1f} GB/s  ({ok})", flush=True)


if __name__ == "__main__":
    main()
