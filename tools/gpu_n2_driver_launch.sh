set -u
O=gpurun_out; mkdir -p $O
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 > $O/fin_n2_ours.json 2> $O/fin_n2_ours.err); tail -2 $O/fin_n2_ours.err
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > $O/fin_n2_ref.json 2> $O/fin_n2_ref.err); cut -c1-200 $O/fin_n2_ref.json
python - <<'PY'
import json
j=json.load(open("gpurun_out/fin_n2_ours.json"))
print("N=2:", j["config"]["rows_total"], "rows; value", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "frac", round(j["roofline"]["frac"],3), "steps", j["steps"], j["clocks"])
PY
