set -u
O=gpurun_out; mkdir -p $O
python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_plain_c3b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:orr_batch_gemm_kernel -s 7 -c 1 -o $O/r02_c3_gemm python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_c3b.log 2>&1
tail -1 $O/ncu_c3b.log
bash tools/gpu_validate.sh
