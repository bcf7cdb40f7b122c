python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/sweep.py --set cplxall --reps 7 --out gpurun_out/cplxall.jsonl > gpurun_out/cplxall.txt 2>&1
python tools/sweep.py --set all --only cplx5,cplx6,sym6,cfg1 --reps 10 --out gpurun_out/sweep_s6.jsonl > gpurun_out/sweep_s6.txt 2>&1
