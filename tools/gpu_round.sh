#!/bin/bash
# One GPU session: bench lines (c2 default, c3, c5, 1-pass variants), then ncu launch lists and full captures.
# Every ncu run follows a plain run of the same command that exited 0.
set -u
O=gpurun_out
mkdir -p $O
python bench.py --steps 200 --warmup 20 > $O/r1_bench_c2.json 2> $O/r1_bench_c2.err || echo "bench c2 FAILED"
tail -c 600 $O/r1_bench_c2.json
python bench.py --workload c3 > $O/r1_bench_c3.json 2> $O/r1_bench_c3.err || { echo "bench c3 FAILED"; tail -5 $O/r1_bench_c3.err; }
python bench.py --workload c5 > $O/r1_bench_c5.json 2> $O/r1_bench_c5.err || { echo "bench c5 FAILED"; tail -5 $O/r1_bench_c5.err; }
python bench.py --workload c3 --batch-passes 1 --no-cpu-baseline > $O/r1_bench_c3_p1.json 2> $O/r1_bench_c3_p1.err || echo "bench c3 p1 FAILED"
python bench.py --workload c5 --batch-passes 1 --no-cpu-baseline > $O/r1_bench_c5_p1.json 2> $O/r1_bench_c5_p1.err || echo "bench c5 p1 FAILED"
for f in c3 c5 c3_p1 c5_p1; do python - <<PY
import json
try:
    j=json.load(open("$O/r1_bench_$f.json"))
    print("$f", round(j["value"]), "QPS dev;", round(j["e2e"]["value"]), "e2e;", round(j["value_warm_terms"]), "warm;", "main ms", round(j["roofline"]["kernel_ms"],3), "frac", round(j["roofline"]["frac"],3), "redo", j["queries_rerun_singly"], j["clocks"])
except Exception as e: print("$f", "unreadable", e)
PY
done
# ncu: launch list of the c3 bench, then one full capture of the main GEMM launch
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $O/plain_c3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1_launches_c3.csv \
    python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:orr_batch_gemm_kernel -s 7 -c 1 -o $O/r1_gemm_c3 -f \
    python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_c3_full.log 2>&1
ncu -i $O/r1_gemm_c3.ncu-rep --page raw --csv > $O/r1_gemm_c3_raw.csv 2>/dev/null
ls -la $O | tail -20
