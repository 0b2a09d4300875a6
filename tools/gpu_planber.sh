#!/bin/bash
for a in maxlog_s16 logmap_s16; do
python tools/plan_ber_parity.py --algo $a --json gpurun_out/r02_plan_ber_parity_$a.json > gpurun_out/planber_$a.log 2>&1
tail -1 gpurun_out/planber_$a.log
done
python -m pytest tests/test_gpu_ber.py -q -k "auto_plan" 2>&1 | tail -3
