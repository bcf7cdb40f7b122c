#!/usr/bin/env python
"""compare sweep jsonl files side by side: python tools/cmp.py a.jsonl b.jsonl ..."""
import json, sys
def load(f):
    d = {}
    for l in open(f):
        r = json.loads(l)
        if not r.get("variant"):
            d[(r["name"], r["dtype"], r["q"])] = r
    return d
runs = [load(f) for f in sys.argv[1:]]
keys = list(runs[-1].keys())
for k in keys:
    r = runs[-1][k]
    cols = " ".join(f"{d[k]['gbs_med']:6.0f}" if k in d and "gbs_med" in d[k] else "     -" for d in runs)
    print(f"{k[0]:7s} {k[1]:4s} q{k[2]} {str(r['view']):26s} k{r['kernel']} v{r['vec']} tx{r['tx']:<3d} ty{r['ty']:<3d} ks{r['ksplit']:<3d}| {cols}")
low = [runs[-1][k]["gbs_med"] for k in keys if "gbs_med" in runs[-1][k]]
import statistics
print("last run: min %.0f  median %.0f  geo-mean %.0f  count>=5243 (0.8 of measured peak): %d/%d" % (min(low), statistics.median(low), statistics.geometric_mean(low), sum(x >= 5243 for x in low), len(low)))
