#!/bin/bash
# gpurun --gpus 8 -- bash tools/gpu_multi.sh 8
# N GPUs: the single-process cluster (single calls, pipelined run, batch) and the row-sharded batched configs under torchrun
set -u
O=gpurun_out; mkdir -p $O
N=${1:-8}
(time python bench.py --cluster --gpus $N --steps 100 --warmup 10 > $O/r2n${N}_cluster2.json 2> $O/r2n${N}_cluster2.err); tail -3 $O/r2n${N}_cluster2.err
python - <<PY
import json
j=json.load(open("gpurun_out/r2n${N}_cluster2.json"))
print("cluster N=$N:", j["config"]["rows_total"], "rows; single calls corpus_qps", round(j["corpus_qps"],2), "call", j["e2e"]["call_ms"], "| pipelined run", j["pipelined_run"], "|", j.get("cluster_batch"))
PY
for W in c3 c5; do
  (time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload $W --steps 10 --warmup 3 > $O/r2n${N}_$W.json 2> $O/r2n${N}_$W.err); tail -2 $O/r2n${N}_$W.err
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2n${N}_$W.json"))
    print("$W N=$N:", j["config"]["rows_total"], "rows; value", round(j["value"]), "corpus_qps", round(j["corpus_qps"]), "ms/step", round(j["ms_per_step"],3), "main ms", round(j["roofline"]["kernel_ms"],3), j["clocks"])
except Exception as e: print("unreadable", e)
PY
done
