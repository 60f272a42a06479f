#!/bin/bash
# first GPU visit: environment, parity tests, gather probe, bench (both arms), ncu launch list + full capture
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv
nproc; free -g | head -2; lscpu | grep -E "Model name|^CPU\(s\)|Socket"
} > gpurun_out/env.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 tools/bin/probe_gather > gpurun_out/probe_gather.txt 2>&1
cat gpurun_out/probe_gather.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err; tail -3 gpurun_out/bench_b200.err; cat gpurun_out/bench_b200.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bv_(rank|select)_kernel' -s 6 -c 2 -o gpurun_out/prof_bv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out
