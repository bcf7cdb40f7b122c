#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list: per-kernel launches, time, share.
    python tools/launch_summary.py gpurun_out/launches.csv "<command that was profiled>" > profiles/rNN_bench_launch_list.txt"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    agg.setdefault(r["Kernel Name"][:100], []).append(float(r["Metric Value"].replace(",", "")) / 1e6)
tot = sum(sum(v) for v in agg.values())
print(sys.argv[2] if len(sys.argv) > 2 else "")
print("(per-launch times are cold-cache and serialised: compare SHARES, not absolutes; fill kernels build the synthetic inputs outside the timed region)\n")
print(f"{'kernel':102s} {'launches':>8s} {'total ms':>10s} {'avg ms':>9s} {'share':>7s}")
for k, v in agg.items():
    print(f"{k:102s} {len(v):8d} {sum(v):10.3f} {sum(v) / len(v):9.4f} {100 * sum(v) / tot:6.1f}%")
ttv = [(k, v) for k, v in agg.items() if "ttv_col" in k or "ttv_dot" in k or "reduce" in k or "ttv_stream" in k]
tt = sum(sum(v) for _, v in ttv) or 1.0
print("\nshare among the TTV kernels of a step:")
for k, v in ttv:
    print(f"  {k:100s} {100 * sum(v) / tt:6.1f}%")
