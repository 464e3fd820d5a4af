#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_text.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/text_check.py > $O/it5_text_check.log 2>&1; tail -12 $O/it5_text_check.log
