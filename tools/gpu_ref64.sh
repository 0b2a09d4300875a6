#!/bin/bash
# GPU round trip for the fp64 reference-order kernel: parity tests, timing, optional ncu capture.
#   gpurun -- tools/gpu_ref64.sh <tag> [ncu]
tag=${1:-x}
timeout 600 python -m pytest tests/test_gpu_ref64.py tests/test_gpu_modem.py tests/test_gpu_ratematch.py -x -q 2>&1 | tail -5
for n in 512 4096; do timeout 120 python tools/time_ref64.py $n 2>&1 | tail -1 | tee -a gpurun_out/ref64_time_$tag.txt; done
if [ "$2" = "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ref64_decode -s 2 -c 1 -o gpurun_out/prof_ref64_$tag -f python tools/time_ref64.py 4096 > gpurun_out/ncu_ref64_$tag.log 2>&1
  tail -2 gpurun_out/ncu_ref64_$tag.log | cut -c1-200
fi
