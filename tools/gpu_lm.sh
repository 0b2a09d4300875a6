#!/bin/bash
# One GPU round trip while tuning the packed-16-bit Log-MAP kernel: parity tests, timings.
#   gpurun -- tools/gpu_lm.sh <tag>
tag=${1:-x}
python -m pytest tests/test_gpu_logmap_s16.py -x -q 2>&1 | tail -15
python tools/time_fast.py 4096 algo=logmap_s16 2>&1 | tail -2
python tools/time_fast.py 4096 algo=logmap_s16 warmup=16 sub_block=48 2>&1 | tail -2
python tools/time_fast.py 4096 algo=maxlog_s16 2>&1 | tail -1
