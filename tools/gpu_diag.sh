#!/bin/bash
mkdir -p gpurun_out
python tools/bench_binned.py --chunks 24 --ops rank1,select1 --reps 7 2>&1 | grep binned | cut -c1-140
python bench.py --steps 20 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read())
print('bench', r['value']/1e9, r['ms_per_step'], {k:v['kernel_ms'] for k,v in r['roofline_by_kernel'].items()}, r['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'bin_' -s 12 -c 12 --csv --log-file gpurun_out/diag_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > /dev/null 2>&1
python tools/summarize_launch_csv.py gpurun_out/diag_launches.csv | head -14
