#!/usr/bin/env python
"""tools/one.py -- run ONE config of tools/sweep.py a few times (for ncu captures).

    ncu --set full --clock-control none --import-source on -k regex:ttv_ -s 3 -c 1 -o gpurun_out/prof \
        python tools/one.py --cfg sym5 --q 1 [--ksplit N] [--kernel colx] [--launches 4]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sweep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", required=True)
    ap.add_argument("--q", type=int, required=True)
    ap.add_argument("--dtype", default="")
    ap.add_argument("--ksplit", type=int, default=0)
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--launches", type=int, default=4)
    ap.add_argument("--set", default="all", help="config family of tools/sweep.py the name comes from (pad, padv, ... are not in 'all')")
    args = ap.parse_args()
    for name, dt, na, pia, q, *rest in sweep.configs(args.set):
        if name == args.cfg and q == args.q and (not args.dtype or dt == args.dtype):
            opts = {}
            if args.ksplit:
                opts["ksplit"] = args.ksplit
            if args.kernel != "auto":
                opts["kernel"] = args.kernel
            r = sweep.bench_one(dt, na, pia, q, reps=max(1, args.launches - 3), wa=rest[0] if rest else None, **opts)
            print(name, dt, q, r)
            return
    raise SystemExit("no such config")


if __name__ == "__main__":
    main()
