#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_eps.py -m gpu -x -q 2>&1 | tail -4
for T in 16 4 0; do
  ORR_BENCH_TERMS=$T ORR_BATCH_TRACE=1 timeout 300 python bench.py --workload c5 --no-cpu-baseline --steps 6 --warmup 3 2>&1 >/dev/null | grep "orr batch" | grep "B=256" | tail -2 | sed "s/^/terms=$T /"
done
ORR_BATCH_TRACE=1 timeout 300 python bench.py --workload c3 --no-cpu-baseline --steps 4 --warmup 3 2>&1 >/dev/null | grep "orr batch" | tail -2
# K3 at 300 rows (c1, cap=300) and the no-embedding kernels: full captures
ncu --set full --clock-control none --import-source on -k regex:orr_rescore_kernel -s 60 -c 1 -o $O/r2_k3_c1 -f python tools/probe_r2.py c1 > $O/ncu_k3.log 2>&1
ncu -i $O/r2_k3_c1.ncu-rep --page raw --csv > $O/r2_k3_c1_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/r2_k3_c1_raw.csv | head -30
ncu -i $O/r2_k3_c1.ncu-rep --page source --csv > $O/r2_k3_c1_src.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:orr_noemb -s 12 -c 1 -o $O/r2_noemb3 -f python tools/probe_r2.py noemb > $O/ncu_noemb3.log 2>&1
ncu -i $O/r2_noemb3.ncu-rep --page raw --csv > $O/r2_noemb3_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/r2_noemb3_raw.csv | head -24
ncu -i $O/r2_noemb3.ncu-rep --page source --csv > $O/r2_noemb3_src.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2_launches_noemb3.csv python tools/probe_r2.py noemb > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2_launches_noemb3.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows: agg[r[4][:60]].append(float(r[-1]))
for k,v in agg.items(): print(f"{k:60s} n={len(v):3d} median={sorted(v)[len(v)//2]:.0f} ns")
PY
