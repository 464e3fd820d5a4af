#!/bin/bash
# N=8: the driver's scaling launch (5M rows/GPU = configs[3], 40M x 3072) and the single-process cluster form
set -u
O=gpurun_out; mkdir -p $O
N=${1:-8}
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 60 --warmup 10 > $O/r2n${N}_ours.json 2> $O/r2n${N}_ours.err); tail -3 $O/r2n${N}_ours.err
python - <<PY
import json
j=json.load(open("gpurun_out/r2n${N}_ours.json"))
print("N=$N:", j["config"]["rows_total"], "rows; value", round(j["value"],1), "corpus_qps", round(j["corpus_qps"],2), "e2e", round(j["e2e"]["value"],1), "frac", round(j["roofline"]["frac"],3), "blocking", round(j["value_blocking_exchange"],1), j["per_rank_scan_kernel_ms"], j["clocks"])
s=j.get("rows_1m_per_gpu")
if s: print("  1M/GPU:", round(s["value"],1), "e2e", round(s["e2e"]["value"],1))
PY
(time python bench.py --cluster --gpus $N --steps 60 --warmup 10 > $O/r2n${N}_cluster.json 2> $O/r2n${N}_cluster.err); tail -3 $O/r2n${N}_cluster.err
python - <<PY
import json
j=json.load(open("gpurun_out/r2n${N}_cluster.json"))
print("cluster N=$N:", j["config"]["rows_total"], "rows; value", round(j["value"],1), "corpus_qps", round(j["corpus_qps"],2), "call", j["e2e"]["call_ms"], "frac", round(j["roofline"]["frac"],3), j.get("cluster_batch"))
PY
