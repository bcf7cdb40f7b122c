set -x
python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "colx" > gpurun_out/colr_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/colr_pytest.log
ENVS="TTV_B200_COLX_WARP=1;TTV_B200_COLX_WARP=2,TTV_B200_KU=8;TTV_B200_COLX_WARP=2,TTV_B200_KU=4"
python tools/sweep.py --set all --only sym7,sym7d --qs 4,7 --reps 10 --envs "$ENVS" --out gpurun_out/colr_probe.jsonl > gpurun_out/colr_probe.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:ttv_ -s 3 -c 1 -o gpurun_out/prof_stream_sym7d_q2 python tools/one.py --cfg sym7d --q 2 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ttv_ -s 3 -c 1 -o gpurun_out/prof_dotf_cplx5_q5 python tools/one.py --cfg cplx5 --q 5 > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ttv_ -s 3 -c 1 -o gpurun_out/prof_col_sym4d_q2 python tools/one.py --cfg sym4d --q 2 > gpurun_out/ncu_c.log 2>&1
