#!/bin/bash
# round 2, GPU session 9: the reworked one-kernel exchange (emulated ranks), the named sweep with the new chooser rules, the
# scatter kernel's DRAM traffic under ncu (peers emulated on one GPU), bench
out=gpurun_out; mkdir -p $out
(timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -k "exchange or scatter or stream or colx") > $out/r02i_pytest.log 2>&1; tail -4 $out/r02i_pytest.log
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "single_kernel_exchange" > $out/r02i_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $out/r02i_memcheck.log
timeout 240 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "colt and float32" > $out/r02i_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 $out/r02i_racecheck.log
timeout 200 python tools/sweep.py --set named --reps 5 --out $out/r02i_sweep_named.jsonl > $out/r02i_sweep_named.txt 2>&1; tail -2 $out/r02i_sweep_named.txt | cut -c1-160
timeout 60 python tools/probe/exchange_emulated.py > $out/r02i_scatter_emulated.txt 2>&1; cat $out/r02i_scatter_emulated.txt
timeout 120 ncu --set full --clock-control none --import-source on -k regex:ttv_col_scatter -s 3 -c 1 -f -o $out/r02i_ncu_scatter_emulated python tools/probe/exchange_emulated.py > $out/r02i_ncu_scatter.log 2>&1
(time timeout 400 python bench.py) > $out/r02i_bench.json 2> $out/r02i_bench.err; echo "bench rc=$?"; cut -c1-160 $out/r02i_bench.json
