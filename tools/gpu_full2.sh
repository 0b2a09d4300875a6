#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -8
python tools/bler_refchannel.py --algo logmap_s16 --frames 32768 --ebn0 0.0 0.1 0.2 0.3 0.4 0.5 0.6 0.7 0.8 0.9 1.0 --out gpurun_out/r02_bler_refchannel_logmap_s16.json 2>&1 | tail -3
