#!/bin/bash
# quick iteration on the binned path: its parity tests, sanitizer (memcheck + racecheck) on small forced-binned batches,
# direct-vs-binned timing, per-kernel ncu times
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bv_gpu.py -m gpu -x -q -k "binned" > gpurun_out/pytest_binned.log 2>&1; tail -3 gpurun_out/pytest_binned.log
if [ -n "$SANITIZE" ]; then
for tool in memcheck racecheck; do
SDSLGPU_BIN_CHUNK_BYTES=4096 timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_binned.py > gpurun_out/sanitize_binned_$tool.log 2>&1; echo "$tool exit $?" | tee -a gpurun_out/sanitize_binned_$tool.log
tail -3 gpurun_out/sanitize_binned_$tool.log
done
fi
timeout 600 python tools/bench_binned.py --chunks ${CHUNKS:-8,16,32} --ops ${OPS:-rank1,select1} --out gpurun_out/bench_binned.jsonl 2>&1 | tail -14
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'bin_' -c 14 --csv --log-file gpurun_out/binned_launches.csv python tools/bench_binned.py --chunks 16 --reps 1 --ops rank1,select1 > gpurun_out/ncu_binned.log 2>&1
python tools/summarize_launch_csv.py gpurun_out/binned_launches.csv > gpurun_out/binned_launches_summary.txt 2>&1; head -16 gpurun_out/binned_launches_summary.txt
