#!/bin/bash
# round 2, GPU session 12 (EIGHT GPUs): bench at N = 8 with the three exchange forms inside the cfg5 leg, the fixed exchange probe,
# multi-GPU parity with the reworked one-kernel exchange
out=gpurun_out; mkdir -p $out
(time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --steps 20 --warmup 5) > $out/r02l_bench_n8.json 2> $out/r02l_bench_n8.err; echo "bench n8 rc=$?"; cut -c1-160 $out/r02l_bench_n8.json; tail -3 $out/r02l_bench_n8.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29642 tools/probe/exchange_probe.py --reps 40 > $out/r02l_exchange_forms.txt 2>&1; grep -A4 "cfg5 q=3" $out/r02l_exchange_forms.txt
(time timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -p no:cacheprovider) > $out/r02l_pytest.log 2>&1; tail -4 $out/r02l_pytest.log
(time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29643 bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e) > $out/r02l_bench_n4.json 2> $out/r02l_bench_n4.err; echo "bench n4 rc=$?"; cut -c1-160 $out/r02l_bench_n4.json
