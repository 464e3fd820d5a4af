#!/bin/bash
# N=2: the driver's scaling launch of both arms, as the driver issues it
set -u
O=gpurun_out; mkdir -p $O
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > $O/r2n2_ref.json 2> $O/r2n2_ref.err); tail -2 $O/r2n2_ref.err; cut -c1-300 $O/r2n2_ref.json
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 10 > $O/r2n2_ours.json 2> $O/r2n2_ours.err); tail -3 $O/r2n2_ours.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r2n2_ours.json"))
print("N=2:", j["config"]["rows_total"], "rows; value", round(j["value"],1), "corpus_qps", round(j["corpus_qps"],2), "e2e", round(j["e2e"]["value"],1), "frac", round(j["roofline"]["frac"],3), "blocking", round(j["value_blocking_exchange"],1), j["per_rank_scan_kernel_ms"], j["clocks"])
s=j.get("rows_1m_per_gpu")
if s: print("  1M/GPU:", round(s["value"],1), "e2e", round(s["e2e"]["value"],1))
PY
python tools/cluster_check.py 2>&1 | tail -5
