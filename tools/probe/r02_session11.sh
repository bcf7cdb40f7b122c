#!/bin/bash
# round 2, GPU session 11: deeper rings of smaller stages for the small-slab STREAM kernel
out=gpurun_out; mkdir -p $out
V="TTV_B200_STREAM_STAGES=4,TTV_B200_STAGE_KB=16;TTV_B200_STREAM_STAGES=4,TTV_B200_STAGE_KB=20;TTV_B200_STREAM_STAGES=4,TTV_B200_STAGE_KB=24;TTV_B200_STREAM_STAGES=5,TTV_B200_STAGE_KB=12;TTV_B200_STREAM_STAGES=5,TTV_B200_STAGE_KB=16;TTV_B200_STREAM_STAGES=5,TTV_B200_STAGE_KB=20;TTV_B200_STREAM_STAGES=4,TTV_B200_STAGE_KB=40;TTV_B200_STREAM_STAGES=4,TTV_B200_STAGE_KB=52;TTV_B200_STREAM_STAGES=5,TTV_B200_STAGE_KB=40"
timeout 150 python tools/sweep.py --set cplxall --only cx6L --qs 5,6 --reps 5 --envs "$V" --out $out/r02k_stages.jsonl > $out/r02k_stages_cx6.txt 2>&1
timeout 150 python tools/sweep.py --set sym --only sym7 --qs 1,2 --reps 5 --envs "$V" --out $out/r02k_stages.jsonl > $out/r02k_stages_sym7.txt 2>&1
timeout 150 python tools/sweep.py --set fp64 --only sym7d,sym5d --qs 1,2 --reps 5 --envs "$V" --out $out/r02k_stages.jsonl > $out/r02k_stages_f64.txt 2>&1
timeout 100 python tools/sweep.py --set asym --only asym6,asym10 --qs 1,2 --reps 5 --envs "TTV_B200_STREAM_STAGES=4,TTV_B200_STAGE_KB=24;TTV_B200_STREAM_STAGES=5,TTV_B200_STAGE_KB=20" --out $out/r02k_stages.jsonl > $out/r02k_stages_asym.txt 2>&1
cat $out/r02k_stages_*.txt | grep -v cx6L.*c128 | cut -c1-200
