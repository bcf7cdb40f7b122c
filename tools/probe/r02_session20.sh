#!/bin/bash
# round 2, GPU session 20 (one B200): the kernels for tiny extents -- COLF (rows that are not whole vectors), DOTP (fibers of
# two elements): parity tests, A/B against the column kernel, one ncu --set full capture each
out=gpurun_out; mkdir -p $out
(timeout 240 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -k "streamk or dotp or colf" -x 2>&1 | tail -5) > $out/r02w_newkernel_tests.log 2>&1; cat $out/r02w_newkernel_tests.log
timeout 300 python tools/probe/tiny_inner.py colf > $out/r02w_colf_ab6.txt 2>&1; cat $out/r02w_colf_ab6.txt
for spec in "asym6 3 colf_long" "asym5n 2 colf_short" "asym10 1 dotp" "asym3n 2 colf_rows2"; do
  set -- $spec
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:"ttv_(colf|dotp)" -s 3 -c 1 -f -o $out/r02w_ncu_$3 \
    python tools/one.py --set named --cfg $1 --q $2 --dtype f32 --launches 4 > $out/r02w_ncu_$3.log 2>&1
  tail -1 $out/r02w_ncu_$3.log | cut -c1-200
done
ls -la $out/*.ncu-rep | tail -5
