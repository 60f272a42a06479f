#!/bin/bash
# re-verification of HEAD on a fresh box: parity tests, smoke, both bench arms, ncu launch list of the bench command
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $? after $(( $(date +%s)-S )) s" >> gpurun_out/pytest_gpu.log
tail -14 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/bench_n1.json
timeout 600 python bench.py --order direct --no-cpu-baseline > gpurun_out/bench_n1_direct.json 2> gpurun_out/bench_n1_direct.err; cut -c1-200 gpurun_out/bench_n1_direct.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-200 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_bench.log 2>&1
python tools/summarize_launch_csv.py gpurun_out/launches_bench.csv 2>&1 | sed -n '/per kernel/,$p'
echo "total $(( $(date +%s)-S )) s"
