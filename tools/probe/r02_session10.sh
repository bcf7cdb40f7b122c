#!/bin/bash
# round 2, GPU session 10: what limits the small-slab STREAM kernel (ncu --set full on two shapes, two stage sizes)
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv > $out/r02j_smi.txt
for spec in "sym7d 2 f64 24" "sym7d 2 f64 56" "sym7 2 f32 36" "sym7d 1 f64 56"; do
  set -- $spec
  TTV_B200_STAGE_KB=$4 timeout 120 ncu --set full --clock-control none --import-source on -k regex:ttv_stream -s 3 -c 1 -f -o $out/r02j_ncu_$1_q$2_kb$4 python tools/one.py --set named --cfg $1 --q $2 --dtype $3 > $out/r02j_ncu_$1_q$2_kb$4.log 2>&1
done
timeout 100 python tools/sweep.py --set fp64 --only sym7d --qs 1,2 --reps 7 --envs "TTV_B200_STAGE_KB=24;TTV_B200_STAGE_KB=36;TTV_B200_STAGE_KB=56;TTV_B200_STREAM_THREADS=512;TTV_B200_STREAM_THREADS=128" --out $out/r02j_sym7d.jsonl 2>&1 | cut -c1-170
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv >> $out/r02j_smi.txt; cat $out/r02j_smi.txt
