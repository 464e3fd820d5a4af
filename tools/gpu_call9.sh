#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_exact.py tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_text.py tests/test_gpu_service.py tests/test_gpu_maintenance.py -m gpu -x -q 2>&1 | tail -6
python tools/probe_r2.py noemb
python tools/probe_r2.py c1
for W in c5 c3; do
  ORR_BATCH_TRACE=1 timeout 300 python bench.py --workload $W --no-cpu-baseline --steps 8 --warmup 3 2> $O/r2_${W}_v9.err > $O/r2_${W}_v9.json
  grep "orr batch" $O/r2_${W}_v9.err | grep "B=" | tail -2
  python - <<PY
import json
try:
    j=json.load(open("$O/r2_${W}_v9.json"))
    print("  $W:", round(j["value"]), "QPS dev;", round(j["e2e"]["value"]), "e2e; main ms", round(j["roofline"]["kernel_ms"],3), "frac", round(j["roofline"]["frac"],3), "step ms", round(j["ms_per_step"],3), j["clocks"])
except Exception as e: print("unreadable", e)
PY
done
(time python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2_bench_default_v9.json 2> $O/r2_bench_default_v9.err); tail -3 $O/r2_bench_default_v9.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r2_bench_default_v9.json"))
print("c2", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "frac", round(j["roofline"]["frac"],3), "service", round(j["e2e_service"]["value"],1), "x", round(j["e2e_service"]["vs_e2e_ms"],3))
for k,v in j["configs"].items():
    print(k, round(v["value"],1), "e2e", round(v["e2e"]["value"],1), "frac", round(v["roofline"]["frac"],3), "kernel_ms", round(v["roofline"]["kernel_ms"],4), "cpu", round(v.get("cpu_baseline",{}).get("value",0),2), v["e2e"].get("call_ms"))
PY
