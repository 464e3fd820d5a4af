#!/bin/bash
# compute-sanitizer --tool racecheck (shared-memory hazards) and memcheck over the concurrency stress test at a small
# shape, plus one small search through every path.  Run on a GPU box: gpurun -- bash tools/racecheck.sh
set -u
O=gpurun_out
mkdir -p $O
export ORR_STRESS_SMALL=1
for tool in racecheck memcheck; do
  timeout 1500 compute-sanitizer --tool $tool --target-processes all --error-exitcode 9 \
    python -m pytest tests/test_gpu_maintenance.py -m gpu -x -q -k "stress" > $O/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/sanitizer_$tool.log | tail -4
done
