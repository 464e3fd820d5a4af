#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -6
python tools/probe_r2.py c5kw 2>&1 | grep -v "^\[orr" | grep "B=256\|B=1024 terms=[048]:"
for W in c5 c3; do
  ORR_BATCH_TRACE=1 timeout 300 python bench.py --workload $W --no-cpu-baseline --steps 8 --warmup 3 2> $O/r2d_${W}.err > $O/r2d_${W}.json
  grep "orr batch" $O/r2d_${W}.err | grep "B=" | tail -2
  python - <<PY
import json
try:
    j=json.load(open("$O/r2d_${W}.json"))
    print("  $W:", round(j["value"]), "QPS dev;", round(j["e2e"]["value"]), "e2e; main ms", round(j["roofline"]["kernel_ms"],3), "frac", round(j["roofline"]["frac"],3), "step ms", round(j["ms_per_step"],3), j["clocks"])
except Exception as e: print("unreadable", e)
PY
done
