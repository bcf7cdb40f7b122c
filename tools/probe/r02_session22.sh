#!/bin/bash
# round 2, GPU sessions 22 and 26 (one B200): evidence pass of the FINAL tree (COLF with its rows-of-two and short-slab forms, DOTP):
# the whole GPU suite, smoke, both bench arms, the launch list of the bench under ncu, every named config standalone (252 products),
# one ncu --set full capture of each new kernel
out=gpurun_out; mkdir -p $out
bash tools/gpu_session.sh r02z
(time timeout 300 python tools/sweep.py --set named --reps 7 --out $out/r02z_sweep_named.jsonl) > $out/r02z_sweep_named.txt 2>&1; tail -4 $out/r02z_sweep_named.txt | cut -c1-160
for spec in "asym6 3 colf_long" "asym5n 2 colf_short" "asym10 1 dotp" "asym3n 2 colf_rows2"; do
  set -- $spec
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:"ttv_(colf|dotp)" -s 3 -c 1 -f -o $out/r02z_ncu_$3 \
    python tools/one.py --set named --cfg $1 --q $2 --dtype f32 --launches 4 > $out/r02z_ncu_$3.log 2>&1
  tail -1 $out/r02z_ncu_$3.log | cut -c1-200
done
