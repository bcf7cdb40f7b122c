python tools/sweep.py --set pad --reps 7 --out gpurun_out/pad.jsonl > gpurun_out/pad.txt 2>&1
python tools/sweep.py --set quick --reps 7 --out gpurun_out/quick.jsonl > gpurun_out/quick.txt 2>&1
