#!/bin/bash
# r01j: --set full capture of the three kernels of the binned rank and select pipelines at HEAD (new select repair logic)
mkdir -p gpurun_out
S=$(date +%s)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'bin_' -s 6 -c 6 -o gpurun_out/r01j_prof_binned python tools/bench_binned.py --chunks 24 --reps 1 --ops rank1,select1 > gpurun_out/ncu_full_binned.log 2>&1
tail -2 gpurun_out/ncu_full_binned.log; ls -la gpurun_out/r01j_prof_binned.ncu-rep
python tools/summarize_ncu.py gpurun_out/r01j_prof_binned.ncu-rep gpurun_out/r01j_ncu_full_binned.txt > /dev/null 2>&1; grep -c "##" gpurun_out/r01j_ncu_full_binned.txt
echo "total $(( $(date +%s)-S )) s"
