#!/bin/bash
# tools/gpu_session.sh -- the standard evidence pass of one GPU call (≈ 4-5 minutes of box time on one B200):
#   gpurun --timeout 600 -- 'bash tools/gpu_session.sh rNN'
# Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/ (tools/launch_summary.py, tools/ncu_summary.py).
tag=${1:-run}
out=gpurun_out
mkdir -p $out
(time timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8) > $out/${tag}_pytest.log 2>&1
tail -3 $out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-160 $out/${tag}_bench.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; cut -c1-160 $out/${tag}_bench_ref.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-sweep --no-cfg5 --no-cpu > $out/${tag}_bench_under_ncu.log 2>&1
timeout 120 python tools/sweep.py --set quick --out $out/${tag}_sweep_quick.jsonl 2>&1 | tail -8
timeout 120 python tools/sweep.py --set pad --qs 2,3,4 --out $out/${tag}_sweep_pad.jsonl 2>&1 | tail -10
timeout 60 python tools/chain_bench.py --shape 64,64,64,64 --out $out/${tag}_chain.jsonl 2>&1 | tail -2 | cut -c1-200
timeout 100 python tools/e2e_probe.py --gib 1 --nq 16 2>&1 | tail -8
