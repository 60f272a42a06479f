#!/bin/bash
for ls in 6 7 8 9 10; do
echo "== log_s $ls"; SDSLGPU_SELECT_LOG_S=$ls python tools/bench_binned.py --chunks 24 --ops select1 --reps 5 --numpy-words 2>&1 | grep -E "binned|direct" | cut -c1-140
done
