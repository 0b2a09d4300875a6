#!/bin/bash
# Round-end evidence run on one B200:  gpurun -- tools/gpu_round.sh r01
# tests, both bench arms, ncu launch list + one full capture of the dominant kernel.
tag=${1:-r01}
o=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $o/${tag}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $o/${tag}_smoke.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > $o/${tag}_clocks.csv &
SMI=$!
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference_arm.json 2> $o/${tag}_bench_ref.err
python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench.err
kill $SMI
tail -c 600 $o/${tag}_bench_n1.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $o/${tag}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_launch.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fast_s16 -s 3 -c 1 -o $o/${tag}_prof_fast -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_full.log 2>&1
tail -2 $o/${tag}_ncu_full.log | cut -c1-160
# block-length x batch sweep (BASELINE configs[3]) and the caller-side kernels
python tools/sweep_throughput.py --out $o/${tag}_sweep_k_batch.jsonl > $o/${tag}_sweep.log 2>&1; tail -1 $o/${tag}_sweep.log | cut -c1-200
python tools/time_modem.py --json $o/${tag}_modem_timing.json > /dev/null 2>&1
python tools/time_ratematch.py --json $o/${tag}_ratematch_timing.json > /dev/null 2>&1
python tools/time_transport.py --json $o/${tag}_transport_timing.json > /dev/null 2>&1
python tools/profile_callers.py > $o/${tag}_plain_callers.log 2>&1 &&
ncu --set full --clock-control none -k regex:"demap|rate_|crc24|modulate|awgn" -c 14 -o $o/${tag}_prof_callers -f python tools/profile_callers.py > $o/${tag}_ncu_callers.log 2>&1
tail -1 $o/${tag}_ncu_callers.log | cut -c1-160
python tools/sweep_all_sizes.py --json $o/${tag}_all_sizes_auto_plan.json > $o/${tag}_all_sizes.log 2>&1; tail -1 $o/${tag}_all_sizes.log | cut -c1-200
