#!/bin/bash
# Iteration session: flat host term map, Harley-Seal keyword counter, bound-gated keyword application.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -5
ORR_BATCH_TRACE=1 timeout 600 python bench.py --workload c3 --no-cpu-baseline > $O/it3_c3.json 2> $O/it3_c3.err || echo "c3 FAILED"
grep "orr batch" $O/it3_c3.err | tail -4
ORR_BATCH_TRACE=1 timeout 600 python bench.py --workload c5 --no-cpu-baseline > $O/it3_c5.json 2> $O/it3_c5.err || echo "c5 FAILED"
grep "orr batch" $O/it3_c5.err | tail -3
for f in c3 c5; do python - <<PY
import json
try:
    j=json.load(open("$O/it3_$f.json"))
    print("$f", round(j["value"]), "dev;", round(j["e2e"]["value"]), "e2e;", j.get("value_warm_terms"), j["ms_per_step"], j["roofline"].get("kernel_ms"), j["roofline"]["frac"], j["clocks"])
except Exception as e: print("$f", "unreadable", e)
PY
done
ncu --set full --clock-control none --import-source on -k regex:orr_batch_gemm_kernel -s 7 -c 1 -o $O/r1_gemm_c5_v2 -f \
    python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_c5_v2.log 2>&1
ncu -i $O/r1_gemm_c5_v2.ncu-rep --page raw --csv > $O/r1_gemm_c5_v2_raw.csv 2>/dev/null
ls -la $O | grep v2
