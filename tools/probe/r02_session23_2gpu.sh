#!/bin/bash
# round 2, GPU session 23 (TWO GPUs): the multi-GPU tests and the bench at N = 2 on the final tree
out=gpurun_out; mkdir -p $out
(time timeout 300 python -m pytest tests/test_multi_gpu.py tests/test_resident_devices_gpu.py -m gpu -q -p no:cacheprovider) > $out/r02z_pytest_2gpu.log 2>&1; grep -E "passed|failed" $out/r02z_pytest_2gpu.log
(time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 20 --warmup 5) > $out/r02z_bench_n2.json 2> $out/r02z_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-200 $out/r02z_bench_n2.json
