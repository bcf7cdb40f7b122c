#!/bin/bash
# round 2, GPU session 8 (TWO GPUs): bench at N = 2 with the three exchange forms inside the cfg5 leg; the fixed exchange probe
out=gpurun_out; mkdir -p $out
(time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 20 --warmup 5) > $out/r02h_bench_n2.json 2> $out/r02h_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-160 $out/r02h_bench_n2.json; tail -3 $out/r02h_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 tools/probe/exchange_probe.py --reps 20 > $out/r02h_exchange_forms.txt 2>&1; grep -A4 "cfg5 q=3" $out/r02h_exchange_forms.txt
