#!/bin/bash
# round-2 tuning session: cluster-of-4 batched GEMM (query multicast) vs pairs, the reworked no-embedding kernel, c1 launch list
set -u
O=gpurun_out; mkdir -p $O
for CL in 4 2; do
  echo "== batch tests, cluster $CL"
  ORR_BATCH_CLUSTER=$CL ORR_BATCH_TRACE=1 timeout 600 python -m pytest tests/test_gpu_batch.py -m gpu -x -q -k "matches_oracle or bitwise or cascade" 2>&1 | grep -E "passed|failed|Error|error|fit;" | sort | uniq -c | tail -6
done
for CL in 4 2; do
  for W in c5 c3; do
    ORR_BATCH_CLUSTER=$CL timeout 600 python bench.py --workload $W --no-cpu-baseline > $O/r2_${W}_cl$CL.json 2> $O/r2_${W}_cl$CL.err || { echo "bench $W cl$CL FAILED"; tail -3 $O/r2_${W}_cl$CL.err; }
    python - <<PY
import json
try:
    j=json.load(open("$O/r2_${W}_cl$CL.json"))
    print("$W cl$CL:", round(j["value"]), "QPS dev;", round(j["e2e"]["value"]), "e2e; main ms", round(j["roofline"]["kernel_ms"],3), "frac", round(j["roofline"]["frac"],3), "step ms", round(j["ms_per_step"],3), "redo", j["queries_rerun_singly"], j["clocks"])
except Exception as e: print("$W cl$CL unreadable", e)
PY
  done
done
ORR_BATCH_TRACE=1 timeout 300 python bench.py --workload c3 --no-cpu-baseline --steps 4 --warmup 3 2>&1 >/dev/null | grep "orr batch" | tail -4
timeout 300 python -m pytest tests/test_gpu_exact.py -m gpu -x -q 2>&1 | tail -3
python tools/probe_r2.py noemb
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r2_launches_c1.csv python tools/probe_r2.py c1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2_launches_c1.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows: agg[r[4][:60]].append(float(r[-1]))
for k,v in agg.items(): print(f"{k:60s} n={len(v):3d} median={sorted(v)[len(v)//2]:.0f} ns")
PY
