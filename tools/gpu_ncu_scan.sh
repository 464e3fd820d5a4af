set -u
O=gpurun_out; mkdir -p $O
python bench.py --steps 2 --warmup 3 --headline-only --no-cpu-baseline > $O/ncu_plain_scan.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:orr_scan_kernel -s 4 -c 1 -o $O/r02_scan python bench.py --steps 2 --warmup 3 --headline-only --no-cpu-baseline > $O/ncu_scan.log 2>&1
tail -2 $O/ncu_scan.log
python tools/probe_r2.py noemb > $O/ncu_plain_noemb.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:orr_noemb_scores_kernel -s 20 -c 1 -o $O/r02_noemb python tools/probe_r2.py noemb > $O/ncu_noemb2.log 2>&1
tail -2 $O/ncu_noemb2.log
