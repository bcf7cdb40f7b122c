#!/bin/bash
# round 2, GPU session 5 (EIGHT GPUs): bench at N = 8 (weak-scaling workload, cfg5 strong scaling, selfcheck), the exchange
# forms of cfg5 q=3 with NVLink byte counters, one host tensor over 1/2/4/8 GPUs of one process, multi-GPU parity
out=gpurun_out; mkdir -p $out
nvidia-smi topo -m > $out/r02e_topo.txt 2>&1
(time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 20 --warmup 5) > $out/r02e_bench_n8.json 2> $out/r02e_bench_n8.err; echo "bench n8 rc=$?"; cut -c1-200 $out/r02e_bench_n8.json; tail -4 $out/r02e_bench_n8.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 tools/probe/exchange_probe.py --reps 20 > $out/r02e_exchange_forms.txt 2>&1; tail -6 $out/r02e_exchange_forms.txt
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29623 tools/probe/exchange_probe.py --ncu-rank0 $out/r02e_ncu_scatter_rank0.csv --reps 3 > $out/r02e_exchange_ncu.log 2>&1; tail -8 $out/r02e_ncu_scatter_rank0.csv | cut -c1-220
timeout 300 python tools/probe/multi_device_probe.py --gib 16 > $out/r02e_multi_device.txt 2>&1; cat $out/r02e_multi_device.txt
(time timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -p no:cacheprovider) > $out/r02e_pytest.log 2>&1; tail -4 $out/r02e_pytest.log
