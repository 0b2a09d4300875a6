#!/bin/bash
# BER evidence of TDB200_ALGO_LOGMAP_S16 (profiles/r02_*): paired with the reference decoder, and the published table
python tools/bler_paired.py --algo logmap_s16 --frames 16384 --ebn0 0.2 0.3 0.4 0.5 --out gpurun_out/r02_bler_paired_logmap_s16_refchannel.json 2>&1 | tail -6
python tools/bler_paired.py --algo logmap_s16 --frames 16384 --ebn0 0.3 0.4 --gaussian --out gpurun_out/r02_bler_paired_logmap_s16_gaussian.json 2>&1 | tail -3
python tools/bler_refchannel.py --algo logmap_s16 --frames 32768 --ebn0 0.0 0.1 0.2 0.3 0.4 0.5 0.6 0.7 0.8 0.9 1.0 --out gpurun_out/r02_bler_refchannel_logmap_s16.json 2>&1 | tail -2
python tools/bler_refchannel.py --algo maxlog_s16 --frames 32768 --ebn0 0.0 0.1 0.2 0.3 0.4 0.5 0.6 0.7 0.8 0.9 1.0 --out gpurun_out/r02_bler_refchannel_maxlog_s16.json 2>&1 | tail -2
