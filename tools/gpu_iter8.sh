#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -5
ORR_BATCH_TRACE=1 timeout 600 python bench.py --workload c3 --no-cpu-baseline > $O/it8_c3.json 2> $O/it8_c3.err || echo "c3 FAILED"
grep "orr batch" $O/it8_c3.err | tail -4
ORR_BATCH_TRACE=1 timeout 600 python bench.py --workload c5 --no-cpu-baseline > $O/it8_c5.json 2> $O/it8_c5.err || echo "c5 FAILED"
grep "orr batch" $O/it8_c5.err | head -8 | tail -4; grep "orr batch" $O/it8_c5.err | tail -3
for f in c3 c5; do python - <<PY
import json
try:
    j=json.load(open("$O/it8_$f.json"))
    print("$f", round(j["value"]), "dev;", round(j["e2e"]["value"]), "e2e;", j.get("value_warm_terms"), j["ms_per_step"], j["roofline"].get("kernel_ms"), j["roofline"]["frac"], j["clocks"])
except Exception as e: print("$f", "unreadable", e)
PY
done
