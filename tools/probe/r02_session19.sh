#!/bin/bash
# round 2, GPU session 19: final table of every named config (236 products) standalone, and the bench line of the final tree
out=gpurun_out; mkdir -p $out
(time timeout 400 python tools/sweep.py --set named --reps 7 --out $out/r02t_sweep_named.jsonl) > $out/r02t_sweep_named.txt 2>&1; tail -4 $out/r02t_sweep_named.txt | cut -c1-160
(time timeout 400 python bench.py) > $out/r02t_bench.json 2> $out/r02t_bench.err; echo "bench rc=$?"; cut -c1-160 $out/r02t_bench.json; tail -3 $out/r02t_bench.err
(time timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -x) > $out/r02t_pytest.log 2>&1; grep -E "passed|failed" $out/r02t_pytest.log
