python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/chain_bench.py --shape 64,64,64,64 --reps 10 --no-ref --out gpurun_out/chain_small2.jsonl > gpurun_out/chain_small2.txt 2>&1
python tools/sweep.py --set cfg1 --reps 20 --out gpurun_out/q.jsonl > gpurun_out/q.txt 2>&1
python tools/sweep.py --set scal --only scal128 --reps 20 --b2b 40 --out gpurun_out/q.jsonl >> gpurun_out/q.txt 2>&1
