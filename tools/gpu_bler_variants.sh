#!/bin/bash
# The 88-cell table of ITTC/result.txt:102-109 for logmap_s16 with longer guards / sub-blocks (how much of the residual is windowing)
for v in "48 32" "96 48" "192 64"; do set -- $v
python tools/bler_refchannel.py --algo logmap_s16 --frames 32768 --sub-block $1 --warmup $2 --ebn0 0.0 0.1 0.2 0.3 0.4 0.5 0.6 0.7 0.8 0.9 1.0 --out gpurun_out/r02_bler_refchannel_logmap_s16_L$1_G$2.json 2>&1 | tail -1
done
