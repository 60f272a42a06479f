#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python tools/tune_select.py > gpurun_out/tune_select.jsonl 2> gpurun_out/tune_select.err; tail -2 gpurun_out/tune_select.err; cat gpurun_out/tune_select.jsonl
