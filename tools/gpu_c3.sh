#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_compressed_gpu.py -m gpu -x -q -k "binned" > gpurun_out/pytest_c3_binned.log 2>&1; tail -4 gpurun_out/pytest_c3_binned.log
timeout 900 python tools/bench_binned_c3.py --out gpurun_out/bench_binned_c3.jsonl 2>&1 | tail -12
