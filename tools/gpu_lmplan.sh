#!/bin/bash
# After re-tuning the Log-MAP plan table: bit-exact tests, all-sizes throughput, block-error parity of the plan for all 188 sizes.
python -m pytest tests/test_gpu_logmap_s16.py -x -q 2>&1 | tail -3
python tools/sweep_all_sizes.py --algo logmap_s16 --json gpurun_out/r02_all_sizes_logmap_s16.json > gpurun_out/r02_all_sizes_logmap_s16.log 2>&1; tail -1 gpurun_out/r02_all_sizes_logmap_s16.log | cut -c1-200
python tools/plan_ber_parity.py --algo logmap_s16 --json gpurun_out/r02_plan_ber_parity_logmap_s16.json > gpurun_out/planber_logmap_s16.log 2>&1
tail -1 gpurun_out/planber_logmap_s16.log
