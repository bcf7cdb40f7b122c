#!/bin/bash
# round 2, GPU session 7: STREAM stage size / partial sums per output on every STREAM-selected named shape; COLR on complex<float>
out=gpurun_out; mkdir -p $out
S="TTV_B200_STAGE_KB=12;TTV_B200_STAGE_KB=16;TTV_B200_STAGE_KB=20;TTV_B200_STAGE_KB=24;TTV_B200_STAGE_KB=28;TTV_B200_STAGE_KB=32;TTV_B200_STAGE_KB=40;TTV_B200_STAGE_KB=48;TTV_B200_STAGE_KB=56;TTV_B200_STAGE_KB=64;TTV_B200_STAGE_KB=72"
run() {  # $1 tag, rest: extra env
  tag=$1; shift
  env "$@" timeout 150 python tools/sweep.py --set cplxall --only cx6L --qs 5,6 --reps 5 --envs "$S" --out $out/r02g_stream_${tag}.jsonl > $out/r02g_stream_${tag}_cx6.txt 2>&1
  env "$@" timeout 150 python tools/sweep.py --set sym --only sym7 --qs 1,2 --reps 5 --envs "$S" --out $out/r02g_stream_${tag}.jsonl > $out/r02g_stream_${tag}_sym7.txt 2>&1
  env "$@" timeout 150 python tools/sweep.py --set fp64 --only sym7d,sym5d --qs 1,2 --reps 5 --envs "$S" --out $out/r02g_stream_${tag}.jsonl > $out/r02g_stream_${tag}_f64.txt 2>&1
  env "$@" timeout 150 python tools/sweep.py --set asym --only asym6,asym10 --qs 1,2 --reps 5 --envs "TTV_B200_STAGE_KB=24;TTV_B200_STAGE_KB=48" --out $out/r02g_stream_${tag}.jsonl > $out/r02g_stream_${tag}_asym.txt 2>&1
}
run kr1 TTV_B200_X=0
run kr4 TTV_B200_LIB=$PWD/ttv_b200/libttv_b200_kr4.so
timeout 100 python tools/sweep.py --set cplxall --only cx6L,cx6R1 --reps 5 --out $out/r02g_cx6_colr.jsonl > $out/r02g_cx6_colr.txt 2>&1; tail -14 $out/r02g_cx6_colr.txt | cut -c1-160
(timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider -x) > $out/r02g_pytest.log 2>&1; tail -4 $out/r02g_pytest.log
