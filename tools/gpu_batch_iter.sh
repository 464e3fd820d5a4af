#!/bin/bash
# Quick iteration on the batched path: parity tests, then the four bench lines.
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_batch.py -x -q 2>&1 | tail -5
for cfg in "c3 0" "c5 0" "c3 3" "c5 3"; do set -- $cfg
  timeout 300 python bench.py --workload $1 --batch-passes $2 --no-cpu-baseline > $O/it_$1_p$2.json 2> $O/it_$1_p$2.err || { echo "$1 p$2 FAILED"; tail -3 $O/it_$1_p$2.err; }
  python - <<PY
import json
try:
    j=json.load(open("$O/it_$1_p$2.json"))
    print("$1 p$2", round(j["value"]), "QPS dev;", round(j["e2e"]["value"]), "e2e;", round(j["value_warm_terms"]), "warm; main ms", round(j["roofline"]["kernel_ms"],3), "useful frac", round(j["roofline"]["frac"],3), "redo", j["queries_rerun_singly"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e: print("$1 p$2", "unreadable", e)
PY
done
