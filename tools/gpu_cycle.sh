#!/bin/bash
# One GPU round trip while tuning the throughput kernel: parity tests, bench line, full ncu capture.
#   gpurun -- tools/gpu_cycle.sh <tag>
tag=${1:-x}
python -m pytest tests/test_gpu_fast_s16.py -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench_$tag.json').readlines()[-1]); print('BENCH', d['value'], d['ms_per_step'], d['ber'])"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fast_s16 -s 3 -c 1 -o gpurun_out/prof_$tag -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log | cut -c1-200
