#!/bin/bash
# one --set full capture (with source) of each kernel of the binned rank / select pipelines
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bin_' -s 6 -c 6 -o gpurun_out/r01d_prof_binned python tools/bench_binned.py --chunks 24 --reps 1 --ops rank1,select1 > gpurun_out/ncu_full_binned.log 2>&1
tail -3 gpurun_out/ncu_full_binned.log; ls -la gpurun_out/*.ncu-rep
