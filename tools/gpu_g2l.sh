#!/bin/bash
python -m pytest tests/test_gpu_fast_s16.py tests/test_gpu_logmap_s16.py -x -q 2>&1 | tail -5
python tools/plan_ber_parity.py --algo maxlog_s16 --json gpurun_out/planber_fix.json --sizes 88 104 136 152 184 232 248 296 328 344 376 424 472 488 40 64 512 6144 2>&1 | tail -20
python tools/time_fast.py 4096 algo=maxlog_s16 2>&1 | tail -1
