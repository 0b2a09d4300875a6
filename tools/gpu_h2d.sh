#!/bin/bash
# host-to-device ceiling of the box for N = 1, 2, 4, 8 concurrent pinned streams (run under gpurun --gpus 8)
N=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  turbo_decoder_cuda_b200/lib/h2d_control --gpus $n --mb 302 --reps 20 --chunks 8
  turbo_decoder_cuda_b200/lib/h2d_control --gpus $n --mb 302 --reps 20 --chunks 8 --d2h-mb 25
  turbo_decoder_cuda_b200/lib/h2d_control --gpus $n --mb 75.5 --reps 40 --chunks 8 --d2h-mb 25
done | tee gpurun_out/r02_h2d_control.jsonl
nvidia-smi topo -m | head -12
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"
free -g | head -2
