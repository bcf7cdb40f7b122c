#!/bin/bash
# round 2, GPU session 3: whole GPU suite (single-kernel exchange emulated, COLT, resident, multi-device), bench with the
# resident e2e path, launch list of the bench command, memcheck of the new code paths, ncu of the strided DOT kernel
out=gpurun_out; mkdir -p $out
(time timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=6) > $out/r02c_pytest.log 2>&1; tail -12 $out/r02c_pytest.log
(time timeout 400 python bench.py) > $out/r02c_bench.json 2> $out/r02c_bench.err; echo "bench rc=$?"; cut -c1-200 $out/r02c_bench.json; tail -3 $out/r02c_bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $out/r02c_bench_ref.json 2> $out/r02c_bench_ref.err; cut -c1-200 $out/r02c_bench_ref.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r02c_launches.csv python bench.py --steps 2 --warmup 1 --no-sweep --no-cfg5 --no-cpu > $out/r02c_bench_under_ncu.log 2>&1
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "single_kernel_exchange or colt or default_stream or fused_scatter" > $out/r02c_memcheck_a.log 2>&1; echo "memcheck a rc=$?"; tail -4 $out/r02c_memcheck_a.log
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_resident_devices_gpu.py -m gpu -q -p no:cacheprovider -x -k "float32 or keeps_error or copy_and" > $out/r02c_memcheck_b.log 2>&1; echo "memcheck b rc=$?"; tail -4 $out/r02c_memcheck_b.log
timeout 120 ncu --set full --clock-control none --import-source on -k regex:ttv_strided_dot -c 1 -f -o $out/r02c_ncu_strided_dot python tools/one.py --set pad --cfg pad1 --q 1 > $out/r02c_ncu_strided_dot.log 2>&1
ls $out | grep r02c
