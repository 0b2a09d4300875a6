#!/bin/bash
# One GPU round trip while tuning the packed-16-bit Log-MAP kernel: parity tests, timings, optional ncu capture.
#   gpurun -- tools/gpu_lm.sh <tag> [ncu]
tag=${1:-x}
python -m pytest tests/test_gpu_logmap_s16.py -x -q 2>&1 | tail -5
python tools/time_fast.py 4096 algo=logmap_s16 2>&1 | tail -1
python tools/time_fast.py 4096 algo=logmap_s16 warmup=16 sub_block=48 2>&1 | tail -1
python tools/time_fast.py 4096 algo=logmap_s16 warmup=32 sub_block=48 2>&1 | tail -1
python tools/time_fast.py 4096 algo=maxlog_s16 2>&1 | tail -1
if [ "$2" = "ncu" ]; then
ncu --set full --clock-control none --import-source on -k regex:fast_s16 -s 3 -c 1 -o gpurun_out/prof_$tag -f python tools/time_fast.py 4096 algo=logmap_s16 > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log | cut -c1-200
fi
