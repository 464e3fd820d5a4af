#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_exact.py tests/test_gpu_text.py tests/test_gpu_maintenance.py tests/test_gpu_service.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_batch.py -m gpu -x -q -k "matches_oracle or bitwise or cascade" 2>&1 | tail -3
python tools/probe_r2.py noemb
python tools/probe_r2.py c1
for CL in 2 4; do
  ORR_BATCH_CLUSTER=$CL ORR_BATCH_TRACE=1 timeout 300 python bench.py --workload c5 --no-cpu-baseline --steps 6 --warmup 3 2>&1 >/dev/null | grep "orr batch" | sort | uniq -c | sort -rn | head -4
  ORR_BATCH_CLUSTER=$CL timeout 600 ncu --set full --clock-control none --import-source on -k regex:orr_batch_gemm_kernel -s 7 -c 1 -o $O/r2_c5_main_cl$CL -f python bench.py --workload c5 --no-cpu-baseline --steps 3 --warmup 3 > $O/ncu_c5_cl$CL.log 2>&1
  ncu -i $O/r2_c5_main_cl$CL.ncu-rep --page raw --csv > $O/r2_c5_main_cl${CL}_raw.csv 2>/dev/null
  python tools/ncu_summary.py $O/r2_c5_main_cl${CL}_raw.csv 2>/dev/null | grep -E "time_duration|tensor|dram__bytes|dram_throughput|lts__t_sector_hit|xbar2l1tex|lts__t_bytes|grid_size|stall" | head -14
done
ORR_BATCH_TRACE=1 timeout 300 python bench.py --workload c3 --no-cpu-baseline --steps 4 --warmup 3 2>&1 >/dev/null | grep "orr batch" | tail -3
