#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -4
ORR_BATCH_TRACE=1 timeout 300 python bench.py --workload c3 --no-cpu-baseline > $O/it9_c3.json 2> $O/it9_c3.err || echo "c3 FAILED"
grep "orr batch" $O/it9_c3.err | tail -3
ORR_BATCH_TRACE=1 timeout 300 python bench.py --workload c5 --no-cpu-baseline > $O/it9_c5.json 2> $O/it9_c5.err || echo "c5 FAILED"
grep "orr batch" $O/it9_c5.err | tail -2
for f in c3 c5; do python - <<PY
import json
try:
    j=json.load(open("$O/it9_$f.json"))
    print("$f", round(j["value"]), "dev;", round(j["e2e"]["value"]), "e2e;", j["ms_per_step"], j["roofline"].get("kernel_ms"), j["roofline"]["frac"], j["queries_rerun_singly"], j["steps_with_bf16x3_cascade"], j["clocks"])
except Exception as e: print("$f", "unreadable", e)
PY
done
