#!/bin/bash
# Iteration session: text kernel + term-bitmap filter + lazy mid plane + c1 line.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_text.py tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/text_check.py > $O/it2_text_check.log 2>&1; tail -12 $O/it2_text_check.log
ORR_BATCH_TRACE=1 timeout 600 python bench.py --workload c3 --no-cpu-baseline > $O/it2_c3.json 2> $O/it2_c3.err || echo "c3 FAILED"
grep "orr batch" $O/it2_c3.err | tail -4
ORR_BATCH_TRACE=1 timeout 600 python bench.py --workload c5 --no-cpu-baseline > $O/it2_c5.json 2> $O/it2_c5.err || echo "c5 FAILED"
grep "orr batch" $O/it2_c5.err | tail -3
timeout 300 python bench.py --workload c1 > $O/it2_c1.json 2> $O/it2_c1.err || { echo "c1 FAILED"; tail -5 $O/it2_c1.err; }
for f in c3 c5 c1; do python - <<PY
import json
try:
    j=json.load(open("$O/it2_$f.json"))
    print("$f", round(j["value"]), "dev;", round(j["e2e"]["value"]), "e2e;", j.get("value_warm_terms"), j["ms_per_step"], j["roofline"].get("kernel_ms"), j["roofline"]["frac"], j.get("all_rows"), j.get("cpu_baseline"))
except Exception as e: print("$f", "unreadable", e)
PY
done
