set -x
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/sweep.py --set all --reps 10 --out gpurun_out/sweep_s3.jsonl > gpurun_out/sweep_s3.txt 2>&1
python tools/sweep.py --set all --only cplx5,cplx6 --qs 5,2 --reps 10 --envs "TTV_B200_USE_STREAM=1;TTV_B200_USE_DOTF=0" --out gpurun_out/stream_c128.jsonl > gpurun_out/stream_c128.txt 2>&1
