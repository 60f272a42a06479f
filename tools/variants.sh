#!/bin/bash
# tools/variants.sh build|run — compile-time variants of the hot kernels as separate libraries (build: here, on the
# CPU box; run: under gpurun), each timed on the same workload by tools/variant_probe.py.  One JSON line per variant.
set -u
cd "$(dirname "$0")/.."
V=sdsl-lite_b200/build/variants
declare -A FLAGS=(
  [base]=""
  [ahead_r0_s0]="-DBIN_RANK_LOOKAHEAD=0 -DBIN_SEL_LOOKAHEAD=0"
  [ahead_r1_s2]="-DBIN_RANK_LOOKAHEAD=1 -DBIN_SEL_LOOKAHEAD=2"
  [ahead_r2_s1]="-DBIN_RANK_LOOKAHEAD=2 -DBIN_SEL_LOOKAHEAD=1"
)
case ${1:-build} in
build)
  mkdir -p $V
  for name in "${!FLAGS[@]}"; do
    objs=""
    for f in sdsl-lite_b200/csrc/*.cu; do
      o=$V/${name}_$(basename ${f%.cu}).o
      if [ $name != base ] && ! grep -q "BIN_RRR_CTAS\|BIN_SEL_CTAS\|BIN_RANK_LOOKAHEAD" $f && [ -f sdsl-lite_b200/build/$(basename ${f%.cu}).o ]; then
        o=sdsl-lite_b200/build/$(basename ${f%.cu}).o   # files the flags cannot touch: reuse the product's objects
      else
        nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC ${FLAGS[$name]} -c $f -o $o 2> $V/${name}_$(basename ${f%.cu}).log || { cat $V/${name}_$(basename ${f%.cu}).log; exit 1; }
      fi
      objs="$objs $o"
    done
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $V/lib_$name.so $objs -cudart static && echo built $V/lib_$name.so
  done ;;
run)
  for lib in $V/lib_*.so; do
    SDSLGPU_LIB=$PWD/$lib timeout 300 python tools/variant_probe.py $(basename $lib) || echo "{\"variant\": \"$lib\", \"error\": true}"
  done ;;
esac
