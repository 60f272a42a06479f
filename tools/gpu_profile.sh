#!/bin/bash
# ncu evidence for the round: launch list of the bench command + full captures of the dominant kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r01_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bv_(rank|select)_kernel' -s 6 -c 2 -o gpurun_out/r01_prof_bv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'wt_rank_level_kernel|fm_count_kernel|rrr_rank_kernel|sd_rank_kernel' -c 12 -o gpurun_out/r01_prof_wt_fm python tools/bench_all.py --configs C3,C4,C5 --densities 0.1 --cpu-densities "" --csa-log2 26 --csa-ref 0 --reps 1 --queries-c3 1e7 > gpurun_out/ncu_full_wtfm.log 2>&1
ls -la gpurun_out | tail -8
