python tools/sweep.py --set asym --only asym8,asym10 --reps 7 --out gpurun_out/asym_hi.jsonl > gpurun_out/asym_hi.txt 2>&1
