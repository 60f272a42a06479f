#!/bin/bash
# re-verification of HEAD on a fresh box: parity tests, smoke, both bench arms
mkdir -p gpurun_out
S=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $? after $(( $(date +%s)-S )) s" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cut -c1-600 gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-600 gpurun_out/bench_ref.json
echo "total $(( $(date +%s)-S )) s"
