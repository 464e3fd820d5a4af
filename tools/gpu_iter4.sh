#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_batch.py -m gpu -x -q -k "term_counts or planted or rows_matching" 2>&1 | tail -5
ORR_BATCH_TRACE=1 timeout 600 python bench.py --workload c3 --no-cpu-baseline > $O/it4_c3.json 2> $O/it4_c3.err || echo "c3 FAILED"
grep "orr batch" $O/it4_c3.err | tail -4
ORR_BATCH_TRACE=1 timeout 600 python bench.py --workload c5 --no-cpu-baseline > $O/it4_c5.json 2> $O/it4_c5.err || echo "c5 FAILED"
grep "orr batch" $O/it4_c5.err | tail -3
for f in c3 c5; do python - <<PY
import json
try:
    j=json.load(open("$O/it4_$f.json"))
    print("$f", round(j["value"]), "dev;", round(j["e2e"]["value"]), "e2e;", j.get("value_warm_terms"), j["ms_per_step"], j["roofline"].get("kernel_ms"), j["roofline"]["frac"], j["clocks"])
except Exception as e: print("$f", "unreadable", e)
PY
done
# text kernel: launch list + full capture of the substring kernel
timeout 300 python tools/text_check.py 300000 3072 > $O/it4_text_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:orr_text_bits_kernel -s 4 -c 1 -o $O/r1_textbits -f \
    python tools/text_check.py 300000 3072 > $O/ncu_text.log 2>&1
ncu -i $O/r1_textbits.ncu-rep --page raw --csv > $O/r1_textbits_raw.csv 2>/dev/null
ncu -i $O/r1_textbits.ncu-rep --page source --csv > $O/r1_textbits_sass.csv 2>/dev/null
python tools/ncu_summary.py $O/r1_textbits_raw.csv $O/r1_textbits_sass.csv 2>&1 | tail -40
