#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/probe_r2.py noemb
python tools/probe_r2.py c1
for T in 1 0; do
  for W in c5 c3; do
    ORR_PLANES_TILED=$T ORR_BATCH_TRACE=1 timeout 300 python bench.py --workload $W --no-cpu-baseline --steps 8 --warmup 3 2> $O/r2_${W}_tiled$T.err > $O/r2_${W}_tiled$T.json
    grep "orr batch" $O/r2_${W}_tiled$T.err | grep "B=" | tail -2 | sed "s/^/tiled=$T $W /"
    python - <<PY
import json
try:
    j=json.load(open("$O/r2_${W}_tiled$T.json"))
    print("  $W tiled=$T:", round(j["value"]), "QPS dev;", round(j["e2e"]["value"]), "e2e; main ms", round(j["roofline"]["kernel_ms"],3), "frac", round(j["roofline"]["frac"],3), "step ms", round(j["ms_per_step"],3), j["clocks"])
except Exception as e: print("unreadable", e)
PY
  done
done
