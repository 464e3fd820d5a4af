#!/bin/bash
# Last session of the round: the full GPU suite, smoke, and the bench lines of the final code.
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/r1g_pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py > $O/r1g_c2.json 2> $O/r1g_c2.err || { echo "bench c2 FAILED"; tail -5 $O/r1g_c2.err; }
python bench.py --impl reference > $O/r1g_ref.json 2> $O/r1g_ref.err || echo "reference arm FAILED"
ORR_BATCH_TRACE=1 python bench.py --workload c3 > $O/r1g_c3.json 2> $O/r1g_c3.err || { echo "bench c3 FAILED"; tail -5 $O/r1g_c3.err; }
ORR_BATCH_TRACE=1 python bench.py --workload c5 > $O/r1g_c5.json 2> $O/r1g_c5.err || { echo "bench c5 FAILED"; tail -5 $O/r1g_c5.err; }
for f in c2 ref c3 c5; do python - <<PY
import json
try:
    j=json.load(open("$O/r1g_$f.json"))
    r=j.get("roofline",{})
    print("$f", round(j["value"],2), "| e2e", round(j["e2e"]["value"],2), "| ms/step", round(j["ms_per_step"],4), "| roofline", r.get("achieved"), r.get("frac"), "| kernel_ms", r.get("kernel_ms"), "| clocks", j.get("clocks"), "| cpu", (j.get("cpu_baseline") or {}).get("value"))
except Exception as e: print("$f", "unreadable", e)
PY
done
grep "orr batch" $O/r1g_c3.err | tail -2; grep "orr batch" $O/r1g_c5.err | tail -2
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $O/plain_c3g.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1g_launches_c3.csv \
    python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_c3g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:orr_batch_term_bits_kernel -s 4 -c 1 -o $O/r1g_tb -f \
    python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_tbg.log 2>&1
ncu -i $O/r1g_tb.ncu-rep --page raw --csv > $O/r1g_tb_raw.csv 2>/dev/null
