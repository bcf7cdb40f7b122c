#!/usr/bin/env python
"""tools/probe/tiny_inner.py -- A/B of the two kernels for tiny extents (round 2, session 20):
  COLF / STREAMK  column GEMV with a tiny odd inner extent under a long contraction ([outer, n_q, inner], inner = 2, 3, 5, 6, 7 ...:
           rows that no 16-byte vector tiles) against the column kernel with lanes along n_q;
  DOTP     fibers of two elements ([outer, 2, 1]) against STREAM (4-byte types) / DOT (8-byte types).
GB/s (train-timed, median) and sampled parity per shape and setting; the environment switches are read per call."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ttv_b200  # noqa: E402
from ttv_b200.measure import Arena, kernel_label, measure_config  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
aa, ac = Arena(int(13.2e9)), Arena(int(6.6e9))
SWITCHES = ("TTV_B200_USE_STREAMK", "TTV_B200_STREAMK_STAGE_KB", "TTV_B200_USE_DOTP", "TTV_B200_DOTP_KU", "TTV_B200_DOTP_CTAS",
            "TTV_B200_USE_COLF", "TTV_B200_COLF_ITEMS_PER_WARP", "TTV_B200_COLF_CTAS", "TTV_B200_COLF_MIN_SLAB_B", "TTV_B200_COLF_PAIR", "TTV_B200_KSPLIT", "TTV_B200_COLF_SHORT", "TTV_B200_COLF_TINY", "TTV_B200_COLF_TINY_CTAS", "TTV_B200_COLF_PAIR_CTAS")


def run(dt, na, q, settings):
    pia = list(range(1, len(na) + 1))
    cells = []
    for env in settings:
        for k in SWITCHES:
            os.environ.pop(k, None)
        os.environ.update(env)
        pl = ttv_b200.plan(q, na, pia, dtype=dt)
        r = measure_config(dt, na, pia, q, reps=5, warmup=2, arena_a=aa, arena_c=ac)
        tag = ",".join(f"{k.replace('TTV_B200_', '')}={v}" for k, v in env.items()) or "default"
        cells.append(f"{tag}: {kernel_label(pl).split()[0].replace('ttv_', '').replace('_kernel', '')}"
                     f"{'/ks' + str(pl['ksplit']) if pl['ksplit'] > 1 else ''} {r['gbs_med']:5.0f}{'' if r['failures'] == 0 else ' FAIL' + str(r['failures'])}")
    print(f"{dt:4s} {str(na):34s} q={q} view=[{pl['outer']}, {pl['nq']}, {pl['inner']}]  " + " | ".join(cells), flush=True)


if which in ("all", "colf", "colfshort"):
    OFF = {"TTV_B200_USE_COLF": "0", "TTV_B200_USE_STREAMK": "0"}
    S = [OFF, {"TTV_B200_USE_COLF": "0", "TTV_B200_USE_STREAMK": "1"}, {}]
    run("f32", [2, 3, 1 << 20, 2, 4, 16], 3, S)                  # the named asym6 q=3: view [128, 2^20, 6]
    run("i32", [2, 3, 1 << 20, 2, 4, 16], 3, S)
    run("f32", [2, 1 << 17, 2, 4, 2, 2, 64], 2, S)               # the named asym7 q=2: rows of two floats
    for dt, na in [] if which == "colfshort" else [("f32", [3, 1 << 20, 64]), ("f32", [5, 1 << 20, 48]), ("f32", [6, 1 << 20, 32]), ("f32", [7, 1 << 19, 64]),
                   ("f64", [3, 1 << 19, 64]), ("f64", [5, 1 << 19, 48]), ("c64", [3, 1 << 19, 64]), ("f32", [9, 1 << 20, 24]), ("f32", [10, 1 << 20, 24]),
                   ("f32", [2, 1 << 20, 128]), ("i32", [2, 1 << 24, 16]), ("f32", [15, 1 << 18, 64]), ("f32", [21, 1 << 16, 256]), ("f64", [21, 1 << 16, 128]),
                   ("f32", [3, 1 << 26]), ("f32", [85, 1 << 16, 64]), ("f64", [63, 1 << 16, 64])]:
        run(dt, na, 2, S)
    T = [OFF, {}, {"TTV_B200_COLF_CTAS": "16"}, {"TTV_B200_COLF_CTAS": "32"}]
    for dt, na in [("f32", [3, 8192, 1 << 14]), ("f32", [5, 4096, 1 << 15]), ("f32", [3, 1024, 1 << 17]), ("f32", [3, 256, 1 << 18]), ("f32", [5, 64, 1 << 20]),
                   ("f32", [2, 512, 1 << 19]), ("f64", [3, 2048, 1 << 15]), ("f32", [2, 128, 1 << 21]), ("f32", [6, 64, 1 << 20]), ("i32", [2, 16, 1 << 24]),
                   ("f32", [3, 32, 1 << 22]), ("c64", [3, 128, 1 << 19]), ("f64", [3, 16, 1 << 22]), ("f32", [6, 1024, 1 << 16]), ("f32", [2, 128, 2, 2, 1 << 21])]:
        run(dt, na, 2, T)                                        # short slabs: one CTA item is a few batches at most

if which in ("all", "pair"):      # rows of two 4-byte elements: the general COLF kernel against ttv_colf2_kernel
    P2 = [{"TTV_B200_COLF_SHORT": "0"}, {"TTV_B200_COLF_PAIR_CTAS": "6"}, {}, {"TTV_B200_COLF_PAIR_CTAS": "64"}, {"TTV_B200_COLF_PAIR_CTAS": "1000000"}]
    run("f32", [2, 1 << 17, 2, 4, 2, 2, 64], 2, P2)              # the named asym7 q=2
    run("f32", [2, 1 << 20, 512], 2, P2)                         # the named asym3n q=2
    run("f32", [2, 128, 2, 2, 1 << 21], 2, P2)                   # the named asym5n q=2
    run("i32", [2, 128, 2, 2, 1 << 21], 2, P2)
    for dt, na in [("f32", [2, 1 << 20, 128]), ("i32", [2, 1 << 24, 16]), ("f32", [2, 512, 1 << 19]), ("f32", [2, 128, 1 << 21]), ("i32", [2, 16, 1 << 24]),
                   ("f32", [2, 32, 1 << 23]), ("f32", [2, 2048, 1 << 17]), ("f32", [2, (1 << 26) + 1])]:
        run(dt, na, 2, P2)

if which in ("all", "tiny"):      # slabs of 1 .. 16 vectors of two-element rows: what took them before against ttv_colf_tiny_kernel
    Y = [{"TTV_B200_COLF_TINY": "0"}, {}, {"TTV_B200_COLF_TINY_CTAS": "8"}, {"TTV_B200_COLF_TINY_CTAS": "32"}]
    for dt, na in [("f32", [2, 2, 1 << 27]), ("f32", [2, 4, 1 << 26]), ("f32", [2, 8, 1 << 25]), ("f32", [2, 16, 1 << 24]), ("f32", [2, 32, 1 << 23]),
                   ("i32", [2, 2, 1 << 27]), ("i32", [2, 16, 1 << 24]), ("f32", [2, 2, 4, 2, 1 << 15, 2, 3, 2, 2, 128])]:
        run(dt, na, 2, Y)

if which in ("all", "dotp"):
    D = [{"TTV_B200_USE_DOTP": "0"}, {}, {"TTV_B200_DOTP_KU": "4"}, {"TTV_B200_DOTP_CTAS": "64"}]
    D1 = [{"TTV_B200_USE_DOTP": "0"}, {}]
    run("f32", [2, 2, 4, 2, 1 << 15, 2, 3, 2, 2, 128], 1, D)     # the named asym10 q=1: view [1610612736, 2, 1]
    run("i32", [2, 2, 4, 2, 1 << 15, 2, 3, 2, 2, 128], 1, D)
    run("f32", [2, 3, 1 << 20, 2, 4, 16], 1, D)                  # the named asym6 q=1: view [402653184, 2, 1]
    run("i32", [2, 3, 1 << 20, 2, 4, 16], 1, D)
    run("f32", [2, (1 << 24) + 1], 1, D)
    run("f64", [2, 1 << 29], 1, D1)
    run("c64", [2, 1 << 28], 1, D1)
