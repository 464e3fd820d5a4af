#!/bin/bash
# gpurun -- bash tools/gpu_validate.sh
# full validation of the tree as the driver will run it: smoke(), the GPU suite, the default bench line, the reference arm
set -u
O=gpurun_out; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(time timeout 1700 python -m pytest tests -m gpu -x -q) > $O/val_pytest_gpu.log 2>&1; tail -4 $O/val_pytest_gpu.log
(time python bench.py --gpus 1 > $O/val_bench_default.json 2> $O/val_bench_default.err); tail -4 $O/val_bench_default.err
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/val_bench_default.json"))
    print("c2", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "frac", round(j["roofline"]["frac"],3), "service", round(j["e2e_service"]["value"],1), "x", round(j["e2e_service"]["vs_e2e_ms"],3))
    for k,v in j["configs"].items():
        print(k, round(v["value"],1), "e2e", round(v["e2e"]["value"],1), "frac", round(v["roofline"]["frac"],3), "kernel_ms", round(v["roofline"]["kernel_ms"],4), "step ms", round(v["ms_per_step"],4), "cpu", round(v.get("cpu_baseline",{}).get("value",0),2), v["e2e"].get("call_ms"))
except Exception as e: print("unreadable", e)
PY
(time python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/val_bench_reference.json 2> $O/val_bench_reference.err); cut -c1-200 $O/val_bench_reference.json
python tools/ingest_check.py 2>&1 | tail -2
