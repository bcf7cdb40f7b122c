#!/bin/bash
# round 2, GPU session 2: COLT parity + A/B against COL, the honest (rotated, train-timed) table of every named config,
# cfg1 variants, refreshed ncu captures
out=gpurun_out; mkdir -p $out
(timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -k "colt or default_stream") > $out/r02b_pytest.log 2>&1; tail -5 $out/r02b_pytest.log
timeout 200 python tools/sweep.py --set named --reps 5 --out $out/r02b_sweep_named.jsonl > $out/r02b_sweep_named.txt 2>&1; tail -3 $out/r02b_sweep_named.txt
E="TTV_B200_USE_COLT=1;TTV_B200_USE_COLT=1,TTV_B200_COLT_STAGES=6;TTV_B200_USE_COLT=1,TTV_B200_COLT_STAGES=3,TTV_B200_COLT_CTAS=2;TTV_B200_USE_COLT=1,TTV_B200_COLT_STAGE_KB=16,TTV_B200_COLT_STAGES=6,TTV_B200_COLT_CTAS=2;TTV_B200_USE_COLT=1,TTV_B200_COLT_STAGE_KB=64,TTV_B200_COLT_STAGES=3"
timeout 200 python tools/sweep.py --set quick --qs 2,3,4 --reps 5 --envs "$E" --out $out/r02b_colt_ab.jsonl > $out/r02b_colt_ab.txt 2>&1; tail -40 $out/r02b_colt_ab.txt
timeout 100 python tools/sweep.py --set cfg1 --reps 7 --envs "TTV_B200_USE_COLT=1,TTV_B200_KSPLIT=2;TTV_B200_USE_COLT=1,TTV_B200_KSPLIT=4;TTV_B200_USE_COLT=1,TTV_B200_KSPLIT=4,TTV_B200_COLT_STAGES=3,TTV_B200_COLT_CTAS=2;TTV_B200_LOADS=8;TTV_B200_KSPLIT=2;TTV_B200_GRID_MULT=2;TTV_B200_GRID_MULT=3" --out $out/r02b_cfg1_variants.jsonl > $out/r02b_cfg1_variants.txt 2>&1; tail -30 $out/r02b_cfg1_variants.txt
for spec in "cfg1 2 f32" "sym7 3 f32" "cplx5 5 c128" "sym2 1 f32" "sym4 3 f32"; do
  set -- $spec
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:ttv_ -s 3 -c 1 -f -o $out/r02b_ncu_$1_q$2 python tools/one.py --set named --cfg $1 --q $2 --dtype $3 > $out/r02b_ncu_$1_q$2.log 2>&1
done
timeout 120 ncu --set full --clock-control none --import-source on -k regex:ttv_colt -s 3 -c 1 -f -o $out/r02b_ncu_colt_sym4_q3 python tools/one.py --set named --cfg sym4 --q 3 --dtype f32 --kernel colt > $out/r02b_ncu_colt.log 2>&1
ls -la $out | tail -20
