#!/bin/bash
# gpurun -- bash tools/gpu_profile.sh   : full GPU suite, default bench line (all configs), reference arm, launch lists.
set -u
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/prof_gpu.txt 2>&1
(time timeout 1700 python -m pytest tests -m gpu -x -q) > $O/prof_pytest_gpu.log 2>&1; tail -8 $O/prof_pytest_gpu.log
(time python bench.py --gpus 1 > $O/prof_bench_default.json 2> $O/prof_bench_default.err); tail -4 $O/prof_bench_default.err
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/prof_bench_default.json"))
    print("c2", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "frac", round(j["roofline"]["frac"],3), "service", round(j["e2e_service"]["value"],1), "x", round(j["e2e_service"]["vs_e2e_ms"],3))
    for k,v in j["configs"].items():
        print(k, round(v["value"],1), "e2e", round(v["e2e"]["value"],1), "frac", round(v["roofline"]["frac"],3), "kernel_ms", round(v["roofline"]["kernel_ms"],4), "step ms", round(v["ms_per_step"],4), "cpu", round(v.get("cpu_baseline",{}).get("value",0),2), v["e2e"].get("call_ms"))
except Exception as e: print("unreadable", e)
PY
(time python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/prof_bench_reference.json 2> $O/prof_bench_reference.err); cat $O/prof_bench_reference.json | cut -c1-600
for W in c5 c3; do
  ORR_BATCH_TRACE=1 timeout 300 python bench.py --workload $W --no-cpu-baseline --steps 8 --warmup 3 2> $O/prof_${W}.err > $O/prof_${W}.json
  grep "orr batch" $O/prof_${W}.err | grep "B=" | tail -2
done
python bench.py --steps 2 --warmup 3 --headline-only --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/prof_launches_c2.csv python bench.py --steps 2 --warmup 3 --headline-only --no-cpu-baseline > $O/prof_ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/prof_launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > $O/prof_ncu_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/prof_launches_c5.csv python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline > $O/prof_ncu_c5.log 2>&1
ls -la $O | tail -20
