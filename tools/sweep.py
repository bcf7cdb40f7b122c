#!/usr/bin/env python
"""tools/sweep.py -- effective HBM GB/s of the TTV kernels over the BASELINE.json config families (device-resident,
CUDA events, inputs larger than L2 or rotated).  Development / evidence tool; writes JSON lines.

    python tools/sweep.py [--set quick|cfg1|sym|asym|complex|fp64|all] [--out gpurun_out/sweep.jsonl] [--variants]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import ttv_b200  # noqa: E402

TORCH_DT = {"f32": torch.float32, "f64": torch.float64, "c64": torch.complex64, "c128": torch.complex128,
            "i32": torch.int32, "i64": torch.int64}
SIZE = {"f32": 4, "f64": 8, "c64": 8, "c128": 16, "i32": 4, "i64": 8}
L2 = 126 * 2 ** 20
B2B = 0          # --b2b N: also time N launches back to back inside one event pair


def configs(which):
    out = []
    first = lambda p: list(range(1, p + 1))
    last = lambda p: list(range(p, 0, -1))
    if which in ("quick", "cfg1", "all"):
        out += [("cfg1", "f32", [512, 512, 512], first(3), q) for q in (1, 2, 3)]
    if which == "scal":      # size series of the cfg1 shape: fixed cost per launch against streaming rate
        out += [("scal%d" % m, "f32", [512, 512, m], first(3), q) for m in (128, 256, 512, 1024, 2048, 4096) for q in (1, 2, 3)]
    if which == "dotk":      # fibers of 1 .. 16 KB at 4 GiB and at 512 MiB: lanes per fiber / CTA size of the DOT kernel
        out += [("dotk%d" % m, "f32", [m, (1 << 30) // m], first(2), 1) for m in (256, 512, 1024, 2048, 4096)]
        out += [("dots%d" % m, "f32", [m, (1 << 27) // m], first(2), 1) for m in (256, 512, 1024, 2048, 4096)]
    if which == "cplxall":   # BASELINE configs[3] in full: order 4..6, complex<float> / complex<double>, last-order + 2 seeded random layouts, every q
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
    if which in ("quick", "sym", "all"):
        out += [("sym4", "f32", [256] * 4, first(4), q) for q in (1, 2, 3, 4)]
    if which in ("sym", "all"):
        out += [("sym2", "f32", [65536, 65536], first(2), q) for q in (1, 2)]
        out += [("sym3", "f32", [1625] * 3, first(3), q) for q in (1, 2, 3)]
        out += [("sym5", "f32", [84] * 5, first(5), q) for q in range(1, 6)]
        out += [("sym6", "f32", [40] * 6, first(6), q) for q in range(1, 7)]
        out += [("sym7", "f32", [23] * 7, first(7), q) for q in range(1, 8)]
    if which in ("fp64", "all"):
        out += [("cfg5/8", "f64", [2048, 2048, 256], first(3), q) for q in (1, 2, 3)]
        out += [("sym2d", "f64", [46340] * 2, first(2), q) for q in (1, 2)]
        out += [("sym3d", "f64", [1290] * 3, first(3), q) for q in (1, 2, 3)]
        out += [("sym4d", "f64", [215] * 4, first(4), q) for q in range(1, 5)]
        out += [("sym5d", "f64", [73] * 5, first(5), q) for q in range(1, 6)]
        out += [("sym6d", "f64", [36] * 6, first(6), q) for q in range(1, 7)]
        out += [("sym7d", "f64", [21] * 7, first(7), q) for q in range(1, 8)]
    if which in ("asym", "all"):
        for dt in ("f32", "i32"):
            out += [("asym5", dt, [4, 1 << 18, 2, 2, 256], first(5), q) for q in (1, 2, 3, 4, 5)]
            out += [("asym4", dt, [16, 1024, 4, 1 << 14], first(4), q) for q in (1, 2, 3, 4)]
            out += [("asym6", dt, [2, 3, 1 << 20, 2, 4, 16], first(6), q) for q in (1, 2, 3, 4, 6)]
            out += [("asym8", dt, [4, 1 << 16, 2, 2, 3, 2, 2, 64], first(8), q) for q in (1, 2, 5, 8)]
            out += [("asym10", dt, [2, 2, 4, 2, 1 << 15, 2, 3, 2, 2, 128], first(10), q) for q in (1, 3, 5, 7, 10)]
    if which in ("complex", "all"):
        out += [("cplx4", "c64", [128] * 4, last(4), q) for q in (1, 2, 4)]
        out += [("cplx4r", "c64", [128] * 4, [3, 1, 4, 2], q) for q in (1, 2, 3, 4)]
        out += [("cplx5", "c128", [40] * 5, last(5), q) for q in (1, 3, 5)]
        out += [("cplx6", "c128", [25, 24, 25, 24, 20, 22], [2, 5, 1, 6, 3, 4], q) for q in (1, 2, 5, 6)]
    return out


def bench_one(dt, na, pia, q, reps=10, wa=None, **opts):
    n = int(np.prod(na, dtype=object))
    s = SIZE[dt]
    copies = max(1, min(4, -(-4 * L2 // (n * s))))         # rotate when A is not much larger than L2
    As = []
    span = n if wa is None else 1 + sum((e - 1) * w for e, w in zip(na, wa))
    for i in range(copies):
        a = torch.empty(span, dtype=TORCH_DT[dt], device="cuda")
        ttv_b200.fill(a, 0x77170001 + i)
        As.append(a)
    b = torch.empty(na[q - 1], dtype=TORCH_DT[dt], device="cuda")
    ttv_b200.fill(b, 0x77170002)
    nc = ttv_b200.generate_output_shape(na, q); pic = ttv_b200.generate_output_layout(pia, q)
    flags = 2 if wa is None else 2 | 8
    wa = ttv_b200.generate_strides(na, pia) if wa is None else list(wa)
    wc = ttv_b200.generate_strides(nc, pic)
    c = torch.empty(n // na[q - 1], dtype=TORCH_DT[dt], device="cuda")
    run = lambda a: ttv_b200.ttv_lowlevel(q, len(na), a, na, wa, pia, b, [na[q - 1]], c, nc, wc, pic, flags=flags, **opts)
    for i in range(3):
        run(As[i % copies])
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for i, (e0, e1) in enumerate(evs):
        e0.record(); run(As[i % copies]); e1.record()
    torch.cuda.synchronize()
    ts = sorted(e0.elapsed_time(e1) for e0, e1 in evs)
    byt = s * (n + na[q - 1] + n // na[q - 1])
    out = {"ms_med": ts[len(ts) // 2], "ms_min": ts[0], "gbs_med": byt / ts[len(ts) // 2] / 1e6, "gbs_best": byt / ts[0] / 1e6, "bytes": byt}
    if B2B:
        # the same launches back to back inside ONE event pair: per-launch time without the event / launch gaps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(B2B):
            run(As[i % copies])
        e1.record()
        torch.cuda.synchronize()
        out["ms_b2b"] = e0.elapsed_time(e1) / B2B
        out["gbs_b2b"] = byt / out["ms_b2b"] / 1e6
    del As
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="quick")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--variants", action="store_true", help="also sweep threads / unroll / ksplit on each config")
    ap.add_argument("--only", default="", help="comma-separated config names to keep")
    ap.add_argument("--envs", default="", help="semicolon-separated env variants, e.g. 'TTV_B200_STREAM=0;TTV_B200_KU=4,TTV_B200_THREADS=128'")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--ksplits", default="", help="comma-separated forced n_q splits to try on each config, e.g. '2,4,8'")
    ap.add_argument("--qs", default="", help="comma-separated modes to keep")
    ap.add_argument("--b2b", type=int, default=0, help="also time N launches back to back inside one event pair")
    args = ap.parse_args()
    global B2B
    B2B = args.b2b
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    peak = 6553.9
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    with open(args.out, "a") as f:
        for name, dt, na, pia, q, *rest in configs(args.set):
            wa = rest[0] if rest else None
            if args.only and name not in args.only.split(","):
                continue
            variants = [dict()]
            for spec in [e for e in args.envs.split(";") if e]:
                variants.append(dict(env=dict(kv.split("=") for kv in spec.split(","))))
            if args.qs and str(q) not in args.qs.split(","):
                continue
            variants += [dict(ksplit=int(k)) for k in args.ksplits.split(",") if k]
            if args.variants:
                variants += [dict(env=dict(TTV_B200_THREADS=t, TTV_B200_KU=ku), ksplit=ks)
                             for t in ("128", "256") for ku in ("4", "8") for ks in (0, 1, 2, 4)]
            for var in variants:
                env = var.get("env", {})
                for k, v in env.items():
                    os.environ[k] = v
                opts = {k: v for k, v in var.items() if k != "env"}
                try:
                    pl = ttv_b200.plan(q, na, pia, dtype=dt, wa=wa, **({**opts, "flags": 8} if wa else opts))
                    r = bench_one(dt, na, pia, q, reps=args.reps, wa=wa, **opts)
                except Exception as exc:
                    r, pl = {"error": str(exc)}, {}
                for k in env:
                    os.environ.pop(k, None)
                rec = {"name": name, "dtype": dt, "na": na, "pia": pia, "q": q, "variant": {**env, **opts},
                       "view": [pl.get("outer"), pl.get("nq"), pl.get("inner")], "kernel": pl.get("kernel"), "vec": pl.get("vec"),
                       "tx": pl.get("tx"), "ty": pl.get("ty"), "ksplit": pl.get("ksplit"), "ctas": pl.get("ctas"), **r}
                if "gbs_med" in r:
                    rec["frac_measured"] = round(r["gbs_med"] / peak, 3)
                    print(f"{name:8s} {dt:5s} q={q} view={rec['view']} k={rec['kernel']} v={rec['vec']} tx={rec['tx']} ty={rec['ty']} "
                          f"ks={rec['ksplit']} nu={pl.get('nu')} ku={pl.get('ku')} ctas={rec['ctas']} {rec['variant']}  {r['ms_med']:.4f} ms  {r['gbs_med']:.0f} GB/s ({rec['frac_measured']:.2f})"
                          + (f"  b2b {r['gbs_b2b']:.0f}" if "gbs_b2b" in r else ""), flush=True)
                else:
                    print(f"{name:8s} {dt:5s} q={q} ERROR {r['error']}", flush=True)
                f.write(json.dumps(rec) + "\n"); f.flush()
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
