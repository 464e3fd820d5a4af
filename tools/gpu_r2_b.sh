#!/bin/bash
# Round-2 GPU session B: cluster-size A/B on the batched configs, ncu --set full of the C5 main pass
set -u
O=gpurun_out; mkdir -p $O
for CLS in 2 4; do
  for W in c5 c3; do
    ORR_BATCH_CLUSTER=$CLS ORR_BATCH_TRACE=1 timeout 300 python bench.py --workload $W --no-cpu-baseline --steps 8 --warmup 3 2> $O/r2b_${W}_cl$CLS.err > $O/r2b_${W}_cl$CLS.json
    grep "orr batch" $O/r2b_${W}_cl$CLS.err | grep "B=" | tail -2 | sed "s/^/cl=$CLS $W /"
  done
done
python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline > $O/r2b_plain_c5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:orr_batch_gemm_kernel -s 7 -c 2 -o $O/r2b_c5_gemm python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline > $O/r2b_ncu_c5.log 2>&1
tail -3 $O/r2b_ncu_c5.log
ORR_BATCH_CLUSTER=4 python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline > $O/r2b_plain_c5_cl4.log 2>&1 && \
ORR_BATCH_CLUSTER=4 ncu --set full --clock-control none --import-source on -k regex:orr_batch_gemm_kernel -s 7 -c 2 -o $O/r2b_c5_gemm_cl4 python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline > $O/r2b_ncu_c5_cl4.log 2>&1
tail -3 $O/r2b_ncu_c5_cl4.log
ls -la $O | grep r2b
