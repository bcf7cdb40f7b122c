#!/bin/bash
# round 2, GPU session 14: why the sweep leg inside bench.py reads lower on the STREAM shapes than the standalone sweep
out=gpurun_out; mkdir -p $out
timeout 300 python bench.py --no-e2e --no-cpu --no-cfg5 > $out/r02n_bench_sweep_first.json 2> $out/r02n_a.err; echo rc=$?
timeout 400 python bench.py > $out/r02n_bench_full.json 2> $out/r02n_b.err; echo rc=$?
timeout 200 python tools/sweep.py --set named --reps 7 --out $out/r02n_sweep_named.jsonl > $out/r02n_sweep_named.txt 2>&1
python - <<'PY'
import json
for f in ("r02n_bench_sweep_first", "r02n_bench_full"):
    d = json.loads(open(f"gpurun_out/{f}.json").read())
    c = d["configs"]; t = {(r[0], r[1], r[2]): r[3] for r in c["table"]}
    print(f, "min", c["min_gbs"], c["min_name"], "median", c["median_gbs"], "below", len(c["below_0p8_nominal"]), "clocks", c.get("clocks"))
    print("   ", {k: t[k] for k in [("sym7d", "f64", 2), ("sym7", "f32", 2), ("sym7", "f32", 3), ("cx6L", "c128", 6), ("cfg1", "f32", 1), ("sym4", "f32", 3)]})
rows = [json.loads(l) for l in open("gpurun_out/r02n_sweep_named.jsonl")]
t = {(r["name"], r["dtype"], r["q"]): round(r["gbs_med"]) for r in rows if "gbs_med" in r}
g = sorted(t.values())
print("standalone min", g[0], "median", g[len(g) // 2], "below", sum(1 for x in g if x < 6400))
print("   ", {k: t[k] for k in [("sym7d", "f64", 2), ("sym7", "f32", 2), ("sym7", "f32", 3), ("cx6L", "c128", 6), ("cfg1", "f32", 1), ("sym4", "f32", 3)]})
PY
