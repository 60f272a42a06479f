#!/bin/bash
# round 1, session 3: egress + device wavelet-tree construction — the tests that exercise them, then build timings
mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python -m pytest tests/test_egress_gpu.py tests/test_wt_gpu.py tests/test_wt_int_gpu.py tests/test_wt_rrr_gpu.py tests/test_load_sdsl_gpu.py \
  "tests/test_fm_gpu.py::test_fm_catalogue" -m gpu -x -q --durations=8 > gpurun_out/pytest_egress.log 2>&1
echo "pytest exit $? after $(( $(date +%s)-S )) s" >> gpurun_out/pytest_egress.log
tail -25 gpurun_out/pytest_egress.log
timeout 400 python tools/bench_build.py 28 ref > gpurun_out/bench_build.jsonl 2> gpurun_out/bench_build.err; tail -3 gpurun_out/bench_build.err; cat gpurun_out/bench_build.jsonl
echo "total $(( $(date +%s)-S )) s"
