#!/bin/bash
# binned batch path: parity tests, memcheck on a small forced-binned run, direct-vs-binned timing, per-kernel times
mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python -m pytest tests/test_bv_gpu.py -m gpu -x -q --durations=5 > gpurun_out/pytest_bv.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_bv.log
tail -12 gpurun_out/pytest_bv.log
echo "t=$(( $(date +%s)-S ))"
SDSLGPU_BIN_CHUNK_BYTES=4096 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_binned.py > gpurun_out/sanitize_binned.log 2>&1; echo "sanitizer exit $?" | tee -a gpurun_out/sanitize_binned.log
tail -4 gpurun_out/sanitize_binned.log
echo "t=$(( $(date +%s)-S ))"
timeout 600 python tools/bench_binned.py --chunks 8,16,32 --out gpurun_out/bench_binned.jsonl 2>&1 | tail -14
echo "t=$(( $(date +%s)-S ))"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'bin_|bv_rank_kernel|bv_select_kernel' -c 24 --csv --log-file gpurun_out/binned_launches.csv python tools/bench_binned.py --chunks 16 --reps 1 --ops rank1,select1 > gpurun_out/ncu_binned.log 2>&1
python tools/summarize_launch_csv.py gpurun_out/binned_launches.csv > gpurun_out/binned_launches_summary.txt 2>&1; tail -34 gpurun_out/binned_launches_summary.txt
echo "t=$(( $(date +%s)-S ))"
