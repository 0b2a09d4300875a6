#!/bin/bash
# local side of a tuning round trip: build (stop on error), CPU tests, then gpurun tools/gpu_cycle.sh <tag>
set -e
tag=${1:-x}
python -m turbo_decoder_cuda_b200.build 2>&1 | grep -E "error|Used" | head -6
python -m turbo_decoder_cuda_b200.build > /dev/null   # non-zero exit stops here if the build is broken
/usr/local/graft/bin/gpurun --timeout 900 -- "tools/gpu_cycle.sh $tag" 2>&1 | grep -E "passed|failed|BENCH|status|Error|error"
python tools/ncu_summary.py gpurun_out/prof_$tag.ncu-rep --src 2>&1 | grep -E "time_duration|pipe_alu|issue_active|stalled_(wait|no_inst|not_sel|math|dispatch|long|barrier|branch|short)|inst_executed.sum"
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > /tmp/t/src_$tag.csv 2>/dev/null
python tools/ncu_regions.py /tmp/t/src_$tag.csv 0.008
