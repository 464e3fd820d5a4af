#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_sharded.py -m gpu -x -q -k "device_resident or cluster_batch" 2>&1 | tail -4
for W in c3 c5; do
  (time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --workload $W --steps 10 --warmup 3 > $O/r2n2_$W.json 2> $O/r2n2_$W.err); tail -3 $O/r2n2_$W.err
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2n2_$W.json"))
    print("$W N=2:", j["config"]["rows_total"], "rows; value", round(j["value"]), "corpus_qps", round(j["corpus_qps"]), "ms/step", round(j["ms_per_step"],3), "main ms", round(j["roofline"]["kernel_ms"],3), j["clocks"])
except Exception as e: print("unreadable", e)
PY
done
