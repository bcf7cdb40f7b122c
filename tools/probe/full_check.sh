python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/sweep.py --set all --reps 10 --out gpurun_out/sweep_s4.jsonl > gpurun_out/sweep_s4.txt 2>&1
