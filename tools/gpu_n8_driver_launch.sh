set -u
O=gpurun_out; mkdir -p $O
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 > $O/fin_n8_ours.json 2> $O/fin_n8_ours.err); tail -2 $O/fin_n8_ours.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/fin_n8_ours.json"))
print("N=8:", j["config"]["rows_total"], "rows; value", round(j["value"],1), "corpus_qps", round(j["corpus_qps"],2), "e2e", round(j["e2e"]["value"],1), "frac", round(j["roofline"]["frac"],3), "blocking", round(j["value_blocking_exchange"],1), "steps", j["steps"], j["clocks"])
s=j.get("rows_1m_per_gpu")
if s: print("  1M/GPU:", round(s["value"],1), "e2e", round(s["e2e"]["value"],1))
PY
