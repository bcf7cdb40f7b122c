#!/bin/bash
# round 2, GPU session 6: chooser variants on the shapes that sit below 0.8 x 8 TB/s (odd extents 25 / 23 / 21 / 73,
# complex<float> / complex<double> / fp32 / fp64): is any existing kernel family better than the chooser's pick?
out=gpurun_out; mkdir -p $out
V="TTV_B200_USE_COLX=0;TTV_B200_COLX_WARP=0;TTV_B200_COLX_WARP=2;TTV_B200_USE_STREAM=0;TTV_B200_USE_STREAM=1;TTV_B200_STREAM_CTAS=1;TTV_B200_STAGE_KB=24;TTV_B200_STAGE_KB=48;TTV_B200_USE_DOTF=0;TTV_B200_USE_DOTF=1;TTV_B200_LOADS=16;TTV_B200_USE_COLX=0,TTV_B200_LOADS=16"
timeout 250 python tools/sweep.py --set cplxall --only cx6L --reps 5 --envs "$V" --out $out/r02f_variants_cx6.jsonl > $out/r02f_variants_cx6.txt 2>&1
timeout 200 python tools/sweep.py --set sym --only sym7 --reps 5 --envs "$V" --out $out/r02f_variants_sym7.jsonl > $out/r02f_variants_sym7.txt 2>&1
timeout 200 python tools/sweep.py --set fp64 --only sym7d,sym5d --reps 5 --envs "$V" --out $out/r02f_variants_f64.jsonl > $out/r02f_variants_f64.txt 2>&1
timeout 100 python tools/sweep.py --set complex --only cplx5 --reps 5 --envs "$V" --out $out/r02f_variants_cplx5.jsonl > $out/r02f_variants_cplx5.txt 2>&1
timeout 100 python tools/sweep.py --set asym --only asym6,asym4 --qs 2,3 --reps 5 --envs "TTV_B200_BDIRECT=0;TTV_B200_KSPLIT=1;TTV_B200_KSPLIT=4;TTV_B200_KSPLIT=20;TTV_B200_LOADS=16" --out $out/r02f_variants_asym.jsonl > $out/r02f_variants_asym.txt 2>&1
wc -l $out/r02f_*.txt
