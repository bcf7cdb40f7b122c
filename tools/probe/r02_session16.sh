#!/bin/bash
# round 2, GPU session 16: shared-memory carve-out pinned for the shared-memory kernels: does the in-bench sweep steady?
out=gpurun_out; mkdir -p $out
for i in 1 2; do
  timeout 300 python bench.py --no-e2e --no-cpu --no-cfg5 > $out/r02p_bench_$i.json 2> $out/r02p_$i.err; echo rc=$?
done
timeout 200 python tools/sweep.py --set named --reps 7 --out $out/r02p_sweep_named.jsonl > $out/r02p_sweep_named.txt 2>&1
python - <<'PY'
import json
keys = [("sym7d", "f64", 2), ("sym7", "f32", 2), ("sym7", "f32", 3), ("cx6L", "c128", 6), ("cx6R2", "c128", 4), ("sym4d", "f64", 1), ("sym4", "f32", 3)]
for i in (1, 2):
    d = json.loads(open(f"gpurun_out/r02p_bench_{i}.json").read())
    c = d["configs"]; t = {(r[0], r[1], r[2]): r[3] for r in c["table"]}
    print("bench", i, "min", c["min_gbs"], c["min_name"], "median", c["median_gbs"], "below", len(c["below_0p8_nominal"]), {k: t[k] for k in keys})
rows = [json.loads(l) for l in open("gpurun_out/r02p_sweep_named.jsonl")]
t = {(r["name"], r["dtype"], r["q"]): round(r["gbs_med"]) for r in rows if "gbs_med" in r}
g = sorted(t.values())
print("standalone min", g[0], "median", g[len(g) // 2], "below", sum(1 for x in g if x < 6400), {k: t[k] for k in keys})
PY
