#!/bin/bash
# tools/variants.sh build|run — compile-time variants of the hot kernels as separate libraries (build: here, on the
# CPU box; run: under gpurun), each timed on the same workload by tools/variant_probe.py.  One JSON line per variant.
set -u
cd "$(dirname "$0")/.."
V=sdsl-lite_b200/build/variants
declare -A FLAGS=(
  [product]=""
  [rrr_walk_only]="-DRRR_SPARSE_PATH=0"
)
case ${1:-build} in
build)
  mkdir -p $V
  for name in "${!FLAGS[@]}"; do
    make -s -j16 -C sdsl-lite_b200 BUILD=build/variants/$name OUT=build/variants/lib_$name.so EXTRA="${FLAGS[$name]}" && echo built $V/lib_$name.so
  done ;;
run)
  for lib in $V/lib_*.so; do
    SDSLGPU_LIB=$PWD/$lib timeout 300 python ${PROBE:-tools/variant_probe.py} $(basename $lib) || echo "{\"variant\": \"$lib\", \"error\": true}"
  done ;;
esac
