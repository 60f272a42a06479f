#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python tools/bench_all.py --out gpurun_out/bench_all_r01b.jsonl > gpurun_out/bench_all_r01b.log 2>&1; tail -2 gpurun_out/bench_all_r01b.log | cut -c1-300
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err; tail -3 gpurun_out/bench_b200.err; cut -c1-400 gpurun_out/bench_b200.json
