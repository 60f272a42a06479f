#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_compressed_gpu.py tests/test_wt_gpu.py tests/test_wt_rrr_gpu.py tests/test_fm_gpu.py -m gpu -x -q -k "binned or wt_huff_catalogue or level_sync or wt_rrr or csa_over_rrr or fm_catalogue" > gpurun_out/pytest_wt_c3.log 2>&1; tail -5 gpurun_out/pytest_wt_c3.log
timeout 600 python tools/bench_binned_c3.py --kinds rrr --densities 0.5,0.125 --out gpurun_out/bench_binned_rrr.jsonl 2>&1 | grep select | cut -c1-260
timeout 900 python tools/bench_all.py --configs C4 --out gpurun_out/bench_c4.jsonl 2>&1 | cut -c1-330 | grep -v '"env"'
