#!/bin/bash
python -m pytest tests/test_gpu_modem.py -x -q 2>&1 | tail -5
python tools/time_modem.py --json gpurun_out/r02_modem_timing.json 2>&1 | tail -12
TDB200_NO_FUSED_DEMAP=1 python tools/time_modem.py 2>&1 | head -6
