#!/bin/bash
# Final session of the round: smoke, the bench lines (c2 headline, reference arm, c1, c3, c5), ncu launch lists of the
# c2 and c3 commands and full captures of the scan, finalize and term-bitmap kernels.  Every ncu run follows a plain
# run of the same command that exited 0.
set -u
O=gpurun_out
mkdir -p $O
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py > $O/r1f_c2.json 2> $O/r1f_c2.err || { echo "bench c2 FAILED"; tail -5 $O/r1f_c2.err; }
python bench.py --impl reference > $O/r1f_ref.json 2> $O/r1f_ref.err || echo "reference arm FAILED"
python bench.py --workload c1 > $O/r1f_c1.json 2> $O/r1f_c1.err || echo "bench c1 FAILED"
ORR_BATCH_TRACE=1 python bench.py --workload c3 > $O/r1f_c3.json 2> $O/r1f_c3.err || { echo "bench c3 FAILED"; tail -5 $O/r1f_c3.err; }
ORR_BATCH_TRACE=1 python bench.py --workload c5 > $O/r1f_c5.json 2> $O/r1f_c5.err || { echo "bench c5 FAILED"; tail -5 $O/r1f_c5.err; }
for f in c2 ref c1 c3 c5; do python - <<PY
import json
try:
    j=json.load(open("$O/r1f_$f.json"))
    r=j.get("roofline",{})
    print("$f", round(j["value"],2), j["unit"][:12], "| e2e", round(j["e2e"]["value"],2), "| ms/step", round(j["ms_per_step"],4), "| roofline", r.get("achieved"), r.get("frac"), "| kernel_ms", r.get("kernel_ms"), "| clocks", j.get("clocks"), "| cpu", (j.get("cpu_baseline") or {}).get("value"))
except Exception as e: print("$f", "unreadable", e)
PY
done
grep "orr batch" $O/r1f_c3.err | tail -2; grep "orr batch" $O/r1f_c5.err | tail -2
# launch lists
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/plain_c2f.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r1f_launches_c2.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/ncu_c2f.log 2>&1
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $O/plain_c3f.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1f_launches_c3.csv \
    python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_c3f.log 2>&1
# full captures
ncu --set full --clock-control none --import-source on -k regex:orr_scan_kernel -s 12 -c 1 -o $O/r1f_scan -f \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/ncu_scanf.log 2>&1
ncu -i $O/r1f_scan.ncu-rep --page raw --csv > $O/r1f_scan_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:orr_batch_finalize_kernel -s 4 -c 1 -o $O/r1f_fin -f \
    python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_finf.log 2>&1
ncu -i $O/r1f_fin.ncu-rep --page raw --csv > $O/r1f_fin_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:orr_batch_term_bits_kernel -s 4 -c 1 -o $O/r1f_tb -f \
    python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_tbf.log 2>&1
ncu -i $O/r1f_tb.ncu-rep --page raw --csv > $O/r1f_tb_raw.csv 2>/dev/null
ls -la $O | grep r1f_
