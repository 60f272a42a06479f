#!/bin/bash
# compile-time variants of bin_apply_kernel (ILP / resident CTAs): same bench, one library each
mkdir -p gpurun_out
cp sdsl-lite_b200/libsdslgpu.so /tmp/lib_default.so
for v in sdsl-lite_b200/variants/*.so; do
  cp $v sdsl-lite_b200/libsdslgpu.so
  echo "== $v"
  timeout 300 python tools/bench_binned.py --chunks 24 --ops rank1,select1 --reps 7 2>&1 | grep binned | cut -c1-140
done | tee gpurun_out/variants.txt
cp /tmp/lib_default.so sdsl-lite_b200/libsdslgpu.so
