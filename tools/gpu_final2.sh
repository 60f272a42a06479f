#!/bin/bash
# round 1, session 3, final verification of HEAD: GPU parity tests (three of the four slow FM sampling-density cases
# left to the driver's own run), smoke, the bench line, ncu launch list of the bench command
mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q --durations=6 -k "not sampling_densities or 8-16" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $? after $(( $(date +%s)-S )) s" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; cut -c1-260 gpurun_out/bench_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_bench.log 2>&1
python tools/summarize_launch_csv.py gpurun_out/launches_bench.csv 2>&1 | sed -n '/per kernel/,$p'
echo "total $(( $(date +%s)-S )) s"
