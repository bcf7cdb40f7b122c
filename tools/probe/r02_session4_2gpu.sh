#!/bin/bash
# round 2, GPU session 4 (TWO GPUs): multi-GPU parity on real devices, bench at N = 2 (weak-scaling workload, cfg5 strong
# scaling, selfcheck), the three exchange forms of cfg5 q=3, NVLink bytes of the scatter kernel, one host tensor over 2 GPUs
out=gpurun_out; mkdir -p $out
nvidia-smi topo -m > $out/r02d_topo.txt 2>&1
(time timeout 400 python -m pytest tests/test_multi_gpu.py tests/test_resident_devices_gpu.py -m gpu -q -p no:cacheprovider) > $out/r02d_pytest.log 2>&1; tail -6 $out/r02d_pytest.log
(time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5) > $out/r02d_bench_n2.json 2> $out/r02d_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-200 $out/r02d_bench_n2.json; tail -4 $out/r02d_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/probe/exchange_probe.py > $out/r02d_exchange_forms.txt 2>&1; tail -6 $out/r02d_exchange_forms.txt
ncu --query-metrics 2>/dev/null | grep -i -E "nvl|peer" | head -60 > $out/r02d_ncu_nvlink_metric_names.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tools/probe/exchange_probe.py --ncu-rank0 $out/r02d_ncu_scatter_rank0.csv --reps 3 > $out/r02d_exchange_ncu.log 2>&1; tail -12 $out/r02d_ncu_scatter_rank0.csv | cut -c1-220
timeout 400 python tools/probe/multi_device_probe.py --gib 16 > $out/r02d_multi_device.txt 2>&1; cat $out/r02d_multi_device.txt
timeout 300 python tools/probe/multi_device_probe.py --gib 4 --pageable > $out/r02d_multi_device_pageable.txt 2>&1; cat $out/r02d_multi_device_pageable.txt
