#!/bin/bash
# round 2, GPU session 13: programmatic dependent launch on/off (trains of short kernels), whole GPU suite with it on, bench
out=gpurun_out; mkdir -p $out
timeout 150 python tools/sweep.py --set quick --reps 7 --envs "TTV_B200_PDL=0" --out $out/r02m_pdl.jsonl > $out/r02m_pdl_quick.txt 2>&1; cut -c1-170 $out/r02m_pdl_quick.txt
timeout 150 python tools/sweep.py --set cplxall --only cx6L,cx4L --reps 5 --envs "TTV_B200_PDL=0" --out $out/r02m_pdl.jsonl > $out/r02m_pdl_cx.txt 2>&1; grep c64 $out/r02m_pdl_cx.txt | cut -c1-170
timeout 100 python tools/sweep.py --set sym --only sym2,sym7 --qs 1,2 --reps 5 --envs "TTV_B200_PDL=0" --out $out/r02m_pdl.jsonl 2>&1 | cut -c1-170
(time timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider -x) > $out/r02m_pytest.log 2>&1; tail -5 $out/r02m_pytest.log
timeout 60 python tools/chain_bench.py --shape 64,64,64,64 --out $out/r02m_chain.jsonl 2>&1 | tail -3 | cut -c1-250
(time timeout 400 python bench.py) > $out/r02m_bench.json 2> $out/r02m_bench.err; echo "bench rc=$?"; cut -c1-160 $out/r02m_bench.json
