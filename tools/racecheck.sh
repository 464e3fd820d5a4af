#!/bin/bash
# compute-sanitizer (ONE tool per GPU call: racecheck = shared-memory hazards, default; or memcheck) over the concurrency
# stress test at a small shape.  Run on a GPU box: gpurun -- bash tools/racecheck.sh [racecheck|memcheck]
set -u
O=gpurun_out
mkdir -p $O
export ORR_STRESS_SMALL=1
tool=${1:-racecheck}
timeout 600 python -m pytest tests/test_gpu_maintenance.py -m gpu -x -q -k "stress" > $O/sanitizer_plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/sanitizer_plain.log; exit 1; }
tail -1 $O/sanitizer_plain.log
timeout 900 compute-sanitizer --tool $tool --target-processes all --error-exitcode 9 \
  python -m pytest tests/test_gpu_maintenance.py -m gpu -x -q -k "stress" > $O/sanitizer_$tool.log 2>&1
echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/sanitizer_$tool.log | tail -4
