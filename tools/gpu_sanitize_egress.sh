#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck; do
timeout 170 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_egress.py > gpurun_out/sanitize_egress_$tool.log 2>&1; echo "$tool exit $?" | tee -a gpurun_out/sanitize_egress_$tool.log
tail -4 gpurun_out/sanitize_egress_$tool.log
done
