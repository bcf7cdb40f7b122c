#!/bin/bash
# round 2, GPU session 18: DOTF with its vectors of b in registers (2 CTAs per SM) against shared memory (3 CTAs per SM)
out=gpurun_out; mkdir -p $out
V="TTV_B200_DOTF_BREG=1"
timeout 200 python tools/sweep.py --set cplxall --only cx4L,cx5L,cx6L --reps 5 --envs "$V" --out $out/r02q_breg.jsonl > $out/r02q_breg_cx.txt 2>&1
timeout 100 python tools/sweep.py --set complex --only cplx5,cplx6 --reps 5 --envs "$V" --out $out/r02q_breg.jsonl > $out/r02q_breg_cplx.txt 2>&1
timeout 100 python tools/sweep.py --set sym --only sym5,sym6 --qs 1 --reps 5 --envs "$V" --out $out/r02q_breg.jsonl > $out/r02q_breg_sym.txt 2>&1
timeout 100 python tools/sweep.py --set asym --only asym4,asym5,asym8 --qs 1 --reps 5 --envs "$V" --out $out/r02q_breg.jsonl > $out/r02q_breg_asym.txt 2>&1
timeout 100 python tools/sweep.py --set fp64 --only sym4d,sym6d --qs 1 --reps 5 --envs "$V" --out $out/r02q_breg.jsonl > $out/r02q_breg_f64.txt 2>&1
python - <<'PY'
import json
rows = {}
for l in open("gpurun_out/r02q_breg.jsonl"):
    r = json.loads(l)
    if "gbs_med" not in r or r["kernel"] != 5: continue
    rows.setdefault((r["name"], r["dtype"], r["q"], tuple(r["view"])), {})["breg" if r["variant"] else "smem"] = (round(r["gbs_med"]), r.get("failures", 0))
for k, d in rows.items(): print(k[0], k[1], "q=%d" % k[2], list(k[3]), d)
PY
