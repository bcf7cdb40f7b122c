#!/usr/bin/env python
"""tools/sweep.py -- effective HBM GB/s of the TTV kernels over the BASELINE.json config families (device-resident,
CUDA events, inputs larger than L2 or rotated).  Development / evidence tool; writes JSON lines.

    python tools/sweep.py [--set quick|cfg1|sym|asym|complex|fp64|all|cplxall|named] [--out gpurun_out/sweep.jsonl] [--variants]

Every product is also CHECKED: sampled outputs of the last launch against a host long-double dot on regenerated fibers
(ttv_b200/selfcheck.py; bit-exact for int32).  The same engine (ttv_b200/measure.py) runs inside bench.py.
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
from ttv_b200.measure import Arena, measure_config  # noqa: E402
from ttv_b200.workloads import configs  # noqa: E402  (the named config families live in the package: bench.py uses them too)

B2B = 0          # --b2b N: also time N launches back to back inside one event pair
_ARENAS = {}


def bench_one(dt, na, pia, q, reps=10, wa=None, check=True, **opts):
    """one product: timing (CUDA events around every launch) + sampled parity of the last launch (ttv_b200.selfcheck)"""
    if not _ARENAS:
        _ARENAS["a"] = Arena(int(17.5e9))
        _ARENAS["c"] = Arena(int(8.8e9))
    out = measure_config(dt, na, pia, q, wa=wa, reps=reps, check=check, arena_a=_ARENAS["a"], arena_c=_ARENAS["c"], **opts)
    if B2B and wa is None:
        # the same launches back to back inside ONE event pair: per-launch time without the event / launch gaps
        a = _ARENAS["a"].buf[: int(np.prod(na, dtype=object)) * ttv_b200.workloads.SIZE[dt]].view(TORCH_DT[dt])
        b = torch.empty(na[q - 1], dtype=TORCH_DT[dt], device="cuda"); ttv_b200.fill(b, 0x77170002)
        nc = ttv_b200.generate_output_shape(na, q); pic = ttv_b200.generate_output_layout(pia, q)
        c = _ARENAS["c"].buf[: (int(np.prod(na, dtype=object)) // na[q - 1]) * ttv_b200.workloads.SIZE[dt]].view(TORCH_DT[dt])
        wa_, wc = ttv_b200.generate_strides(na, pia), ttv_b200.generate_strides(nc, pic)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(B2B):
            ttv_b200.ttv_lowlevel(q, len(na), a, na, wa_, pia, b, [na[q - 1]], c, nc, wc, pic, flags=2, **opts)
        e1.record()
        torch.cuda.synchronize()
        out["ms_b2b"] = e0.elapsed_time(e1) / B2B
        out["gbs_b2b"] = out["bytes"] / out["ms_b2b"] / 1e6
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
    ap.add_argument("--no-check", action="store_true", help="skip the sampled parity check of every product")
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
                    r = bench_one(dt, na, pia, q, reps=args.reps, wa=wa, check=not args.no_check, **opts)
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
                    if r.get("failures"):
                        print(f"   PARITY FAILURE: {r['failures']} of {r['checked']} sampled outputs, worst err/tol {r['worst_err_over_tol']:.3g}", flush=True)
                else:
                    print(f"{name:8s} {dt:5s} q={q} ERROR {r['error']}", flush=True)
                f.write(json.dumps(rec) + "\n"); f.flush()
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
