#!/bin/bash
# FER of the whole device-resident chain (encode -> rate match -> map -> AWGN -> max-log demap -> de-rate-match ->
# decode, early termination on) over modulation x code rate x Eb/N0, with the C++ caller of the C ABI:
#   gpurun -- tools/chain_sweep.sh > profiles/r01_chain_fer.jsonl      (K = 6144, 16384 codeblocks per point)
B=./turbo_decoder_cuda_b200/lib/tdb200_burst
N=${N:-16384}
run() { $B --total $N --gpus 1 --early-term 1 --modulation $1 --E $2 --ebn0 $3; }
for e in 0.4 0.6 0.8 1.0 1.2; do run 1 18444 $e; run 2 18444 $e; done
for e in 1.0 1.4 1.8 2.2 2.6; do run 4 18444 $e; done
for e in 3.0 3.5 4.0 4.5 5.0; do run 6 18444 $e; done
for e in 0.8 1.2 1.6 2.0; do run 2 12288 $e; done          # rate 1/2
for e in 2.0 2.5 3.0 3.5; do run 2 8192 $e; done           # rate 3/4
for e in 2.5 3.0 3.5 4.0; do run 4 12288 $e; done          # 16QAM rate 1/2
for e in 4.5 5.0 5.5 6.0 6.5; do run 6 12288 $e; done      # 64QAM rate 1/2
for e in 7.0 7.5 8.0 8.5 9.0; do run 6 8190 $e; done       # 64QAM rate 3/4
for e in 0.4 0.6 0.8 1.0; do run 2 30720 $e; done          # rate 1/5 (repetition)
