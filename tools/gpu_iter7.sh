#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
ORR_BATCH_TRACE=1 timeout 600 python bench.py --workload c3 --no-cpu-baseline > $O/it7_c3.json 2> $O/it7_c3.err || echo "c3 FAILED"
grep "orr batch" $O/it7_c3.err | tail -4
ORR_BATCH_TRACE=1 timeout 600 python bench.py --workload c5 --no-cpu-baseline > $O/it7_c5.json 2> $O/it7_c5.err || echo "c5 FAILED"
grep "orr batch" $O/it7_c5.err | tail -3
for f in c3 c5; do python - <<PY
import json
try:
    j=json.load(open("$O/it7_$f.json"))
    print("$f", round(j["value"]), "dev;", round(j["e2e"]["value"]), "e2e;", j.get("value_warm_terms"), j["ms_per_step"], j["roofline"].get("kernel_ms"), j["roofline"]["frac"], j["clocks"])
except Exception as e: print("$f", "unreadable", e)
PY
done
python - <<'PY'
# no-embedding mode (the reference's default configuration): exact path timing at 1M rows
import sys, numpy as np
sys.path.insert(0, ".")
import omni_recall_rag_b200 as orr
from omni_recall_rag_b200 import synth
spec = synth.make_spec(3072)
sh = orr.RecallShard(3072, 1_000_000)
sh.fill_synthetic(spec, 0, 1_000_000)
for qi in range(4):
    q = synth.query_host(spec, qi, 1_000_000, n_terms=4)
    sh.search(None, q.terms, spec.now_ticks, 10)
    print("no-embedding query", qi, sh.last_timing())
PY
