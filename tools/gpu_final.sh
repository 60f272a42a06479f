#!/bin/bash
# evidence for the round: all BASELINE configs (both batch orders), full ncu captures of the pipeline kernels
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python tools/bench_all.py --out gpurun_out/r01e_bench_all.jsonl > gpurun_out/bench_all.log 2>&1; tail -2 gpurun_out/bench_all.log | cut -c1-200
echo "t=$(( $(date +%s)-S ))"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bin_' -s 3 -c 6 -o gpurun_out/r01e_prof_binned python tools/bench_binned.py --chunks 24 --reps 1 --ops rank1,select1 > gpurun_out/ncu_full_binned.log 2>&1
python tools/summarize_ncu.py gpurun_out/r01e_prof_binned.ncu-rep gpurun_out/r01e_ncu_full_binned.txt > /dev/null 2>&1; grep -c "##" gpurun_out/r01e_ncu_full_binned.txt
echo "t=$(( $(date +%s)-S ))"
