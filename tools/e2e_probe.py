#!/usr/bin/env python
"""tools/e2e_probe.py -- host-pointer path of the C-ABI on PAGEABLE (numpy) and PINNED host memory: GB/s of A across PCIe.
    python tools/e2e_probe.py [--gib 4]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import ttv_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=4.0)
    ap.add_argument("--nq", type=int, default=256, help="extent of the contracted mode (power of two <= 4096): C is 1/nq of A")
    args = ap.parse_args()
    n4 = max(2, int(args.gib * (1 << 30) / 4 / (1 << 20)))                 # slabs of 4 MiB along the slowest mode
    na, pia = [256, args.nq, (1 << 20) // (256 * args.nq), n4], [1, 2, 3, 4]
    n = int(np.prod(na))
    q = 2
    nc = ttv_b200.generate_output_shape(na, q); pic = ttv_b200.generate_output_layout(pia, q)
    wa = ttv_b200.generate_strides(na, pia); wc = ttv_b200.generate_strides(nc, pic)
    b = np.ones(na[q - 1], np.float32)
    for kind in ("pageable", "pinned"):
        if kind == "pageable":
            a = np.ones(n, np.float32)
            c = np.empty(n // na[q - 1], np.float32)
        else:
            ta = torch.ones(n, dtype=torch.float32).pin_memory(); a = ta.numpy()
            tc = torch.empty(n // na[q - 1], dtype=torch.float32).pin_memory(); c = tc.numpy()
        for mode in ("default", "TTV_B200_BOUNCE_ADAPT=0", "TTV_B200_BOUNCE_OUT=0", "TTV_B200_H2D_CHUNK_MB=0"):
            for key in ("TTV_B200_H2D_CHUNK_MB", "TTV_B200_BOUNCE_ADAPT", "TTV_B200_BOUNCE_OUT"):
                os.environ.pop(key, None)
            if mode != "default":
                key, val = mode.split("=")
                os.environ[key] = val
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                ttv_b200.ttv_lowlevel(q, 4, a, na, wa, pia, b, [len(b)], c, nc, wc, pic)
                ts.append(time.perf_counter() - t0)
            assert abs(float(c[0]) - float(args.nq)) < 1e-3 and abs(float(c[-1]) - float(args.nq)) < 1e-3 and float(c.min()) == float(c.max())
            print(f"{kind:9s} {mode:26s} {n * 4 / min(ts) / 1e9:7.2f} GB/s of A  (best of 3, {min(ts) * 1e3:.1f} ms; C = A/{args.nq})", flush=True)


if __name__ == "__main__":
    main()
