python -m pytest tests/test_ttvpy_gpu.py -x -q -m gpu > gpurun_out/ttvpy_test.log 2>&1; echo rc=$? >> gpurun_out/ttvpy_test.log
python tools/chain_bench.py --shape 64,64,64,64 --reps 10 --no-ref --out gpurun_out/chain_small.jsonl > gpurun_out/chain_small.txt 2>&1
python tools/chain_bench.py --shape 16,16,16,16,16,16 --reps 10 --no-ref --out gpurun_out/chain_small.jsonl >> gpurun_out/chain_small.txt 2>&1
python tools/chain_bench.py --shape 256,256,256,128 --reps 5 --no-ref --out gpurun_out/chain_small.jsonl >> gpurun_out/chain_small.txt 2>&1
