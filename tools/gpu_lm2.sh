#!/bin/bash
python -m pytest tests/test_gpu_logmap_s16.py tests/test_gpu_fast_s16.py tests/test_gpu_ref64.py -x -q 2>&1 | tail -5
python -m pytest tests/test_gpu_ber.py -x -q -k "paired" 2>&1 | tail -5
python tools/bler_paired.py --algo logmap_s16 maxlog_s16 logmap_f64 --frames 2048 --ebn0 0.3 --gaussian --out gpurun_out/paired_small.json 2>&1 | tail -5
nproc
