python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python tools/sweep.py --set all --reps 10 --out gpurun_out/sweep_s3.jsonl > gpurun_out/sweep_s3.txt 2>&1
python tools/chain_bench.py --shape 256,256,256,128 --reps 5 --out gpurun_out/chain.jsonl > gpurun_out/chain.txt 2>&1
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches_s3.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ttv_col_kernel -s 3 -c 1 -o gpurun_out/prof_cfg1_q3_deep python tools/one.py --cfg cfg1 --q 3 > gpurun_out/ncu_d.log 2>&1
