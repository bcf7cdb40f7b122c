"""The synthetic workloads BASELINE.json names (configs[0..4]), as data: (name, dtype, na, pia, q[, wa]).

bench.py (the sweep leg and the parity summary it prints), tools/sweep.py and tests/test_full_size_gpu.py all take their
shapes from here, so that "every named config" means the same list everywhere.  Shapes follow SURVEY.md 8(d): symmetric
tensors of ~2^32 fp32 / ~2^31 fp64 elements for orders 2..7, asymmetric ones with tiny leading extents and small / large
n_q for fp32 and int32 (bit-exact), complex<float> / complex<double> with last-order and seeded random layouts, and
the per-GPU slabs of the 2048^3 fp64 strong-scaling case.
"""
from __future__ import annotations

import numpy as np

SIZE = {"f32": 4, "f64": 8, "c64": 8, "c128": 16, "i32": 4, "i64": 8}
SEED_A, SEED_B = 0x77170001, 0x77170002


def first_order(p):
    return list(range(1, p + 1))


def last_order(p):
    return list(range(p, 0, -1))


def configs(which: str):
    out = []
    first, last = first_order, last_order
    if which in ("quick", "cfg1", "all", "named"):
        out += [("cfg1", "f32", [512, 512, 512], first(3), q) for q in (1, 2, 3)]
    if which == "scal":      # size series of the cfg1 shape: fixed cost per launch against streaming rate
        out += [("scal%d" % m, "f32", [512, 512, m], first(3), q) for m in (128, 256, 512, 1024, 2048, 4096) for q in (1, 2, 3)]
    if which == "dotk":      # fibers of 1 .. 16 KB at 4 GiB and at 512 MiB: lanes per fiber / CTA size of the DOT kernel
        out += [("dotk%d" % m, "f32", [m, (1 << 30) // m], first(2), 1) for m in (256, 512, 1024, 2048, 4096)]
        out += [("dots%d" % m, "f32", [m, (1 << 27) // m], first(2), 1) for m in (256, 512, 1024, 2048, 4096)]
    if which in ("cplxall", "named"):   # BASELINE configs[3] in full: order 4..6, complex<float> / complex<double>, last-order + 2 seeded random layouts, every q
        for pp, ext in ((4, 128), (5, 48), (6, 25)):
            lays = [("L", last(pp))]
            for seed in (1, 2):
                perm = [int(x) + 1 for x in np.random.default_rng(seed).permutation(pp)]
                lays.append(("R%d" % seed, perm))
            for dt in ("c64", "c128"):
                for tag, pia in lays:
                    out += [("cx%d%s" % (pp, tag), dt, [ext] * pp, pia, q) for q in range(1, pp + 1)]
    if which == "pad":       # slices of a packed 256^4 tensor, read in place through wa (TTV_B200_FLAG_HONOR_STRIDES)
        w4 = [1, 256, 256 ** 2, 256 ** 3]
        out += [("pad3", "f32", [256, 256, 250, 256], first(4), q, w4) for q in (1, 2, 3, 4)]      # A[:, :, :250, :]
        out += [("pad1", "f32", [250, 256, 256, 256], first(4), q, w4) for q in (1, 2, 3, 4)]      # A[:250]: padded rows
        out += [("pad12", "f64", [120, 250, 128, 128], first(4), q, [1, 128, 128 * 256, 128 * 256 * 128]) for q in (1, 2, 3, 4)]
    if which == "padv":      # what decides between the vector and the thread-per-output form of the general-stride kernel
        w4 = [1, 256, 256 ** 2, 256 ** 3]
        out += [("pad1b", "f32", [248, 256, 256, 256], first(4), q, w4) for q in (2, 3, 4)]       # rows of 248 of 256 floats
        out += [("pad12L", "f64", [120, 250, 256, 256], first(4), q, [1, 128, 128 * 256, 128 * 256 * 256]) for q in (2, 3, 4)]
        out += [("pad3h", "f32", [256, 256, 250, 32], first(4), q, w4) for q in (2, 3, 4)]         # 2 GB: fewer waves
        out += [("pad3c", "c64", [128, 256, 250, 128], first(4), q, [1, 128, 128 * 256, 128 * 256 * 256]) for q in (2, 3, 4)]
    if which in ("quick", "sym", "all", "named"):
        out += [("sym4", "f32", [256] * 4, first(4), q) for q in (1, 2, 3, 4)]
    if which in ("sym", "all", "named"):
        out += [("sym2", "f32", [65536, 65536], first(2), q) for q in (1, 2)]
        out += [("sym3", "f32", [1625] * 3, first(3), q) for q in (1, 2, 3)]
        out += [("sym5", "f32", [84] * 5, first(5), q) for q in range(1, 6)]
        out += [("sym6", "f32", [40] * 6, first(6), q) for q in range(1, 7)]
        out += [("sym7", "f32", [23] * 7, first(7), q) for q in range(1, 8)]
    if which in ("fp64", "all", "named"):
        out += [("cfg5/8", "f64", [2048, 2048, 256], first(3), q) for q in (1, 2, 3)]
        out += [("sym2d", "f64", [46340] * 2, first(2), q) for q in (1, 2)]
        out += [("sym3d", "f64", [1290] * 3, first(3), q) for q in (1, 2, 3)]
        out += [("sym4d", "f64", [215] * 4, first(4), q) for q in range(1, 5)]
        out += [("sym5d", "f64", [73] * 5, first(5), q) for q in range(1, 6)]
        out += [("sym6d", "f64", [36] * 6, first(6), q) for q in range(1, 7)]
        out += [("sym7d", "f64", [21] * 7, first(7), q) for q in range(1, 8)]
    if which in ("asym", "all", "named"):
        for dt in ("f32", "i32"):
            out += [("asym5", dt, [4, 1 << 18, 2, 2, 256], first(5), q) for q in (1, 2, 3, 4, 5)]
            out += [("asym4", dt, [16, 1024, 4, 1 << 14], first(4), q) for q in (1, 2, 3, 4)]
            out += [("asym6", dt, [2, 3, 1 << 20, 2, 4, 16], first(6), q) for q in (1, 2, 3, 4, 6)]
            out += [("asym8", dt, [4, 1 << 16, 2, 2, 3, 2, 2, 64], first(8), q) for q in (1, 2, 5, 8)]
            out += [("asym10", dt, [2, 2, 4, 2, 1 << 15, 2, 3, 2, 2, 128], first(10), q) for q in (1, 2, 3, 5, 7, 10)]
    if which in ("asym2", "named"):      # the remaining orders of BASELINE configs[2] (p = 2..10): 2, 3, 7, 9 (round 2)
        for dt in ("f32", "i32"):
            out += [("asym2", dt, [4, 1 << 26], first(2), q) for q in (1, 2)]
            out += [("asym3", dt, [16, 1 << 12, 1 << 14], first(3), q) for q in (1, 2, 3)]
            out += [("asym7", dt, [2, 1 << 17, 2, 4, 2, 2, 64], first(7), q) for q in (1, 2, 4, 7)]
            out += [("asym9", dt, [2, 2, 1 << 14, 2, 2, 3, 2, 2, 256], first(9), q) for q in (1, 3, 6, 9)]
            # leading extent 2 (rows of two elements when a later mode is contracted): large / small mode in second place
            out += [("asym3n", dt, [2, 1 << 20, 512], first(3), q) for q in (1, 2, 3)]
            out += [("asym5n", dt, [2, 128, 2, 2, 1 << 21], first(5), q) for q in (1, 2, 3, 4, 5)]
    if which in ("complex", "all", "named"):
        out += [("cplx4", "c64", [128] * 4, last(4), q) for q in (1, 2, 4)]
        out += [("cplx4r", "c64", [128] * 4, [3, 1, 4, 2], q) for q in (1, 2, 3, 4)]
        out += [("cplx5", "c128", [40] * 5, last(5), q) for q in (1, 3, 5)]
        out += [("cplx6", "c128", [25, 24, 25, 24, 20, 22], [2, 5, 1, 6, 3, 4], q) for q in (1, 2, 5, 6)]
    return out


def algo_bytes(dtype: str, na, q: int) -> int:
    """sizeof(T) * (N + n_q + N / n_q): read A once, read b once, write C once"""
    n = int(np.prod(na, dtype=object))
    return SIZE[dtype] * (n + int(na[q - 1]) + n // int(na[q - 1]))
