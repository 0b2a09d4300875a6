#!/bin/bash
# Round-2 evidence run on one B200:  gpurun -- tools/gpu_round2.sh r02
# tests, smoke, both bench arms, ncu launch list, full captures of the three decode kernels, fp64 kernel timing.
tag=${1:-r02}
o=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $o/${tag}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $o/${tag}_smoke.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > $o/${tag}_clocks.csv &
SMI=$!
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference_arm.json 2> $o/${tag}_bench_ref.err
python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench.err
kill $SMI
tail -c 900 $o/${tag}_bench_n1.json
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-early-term"
$B > $o/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $o/${tag}_bench_launches.csv $B > $o/${tag}_ncu_launch.log 2>&1
# the max-log throughput kernel (4th launch = first timed step), the Log-MAP variant (launches 6.. of the same template), the fp64 kernel
ncu --set full --clock-control none --import-source on -k regex:fast_s16 -s 3 -c 1 -o $o/${tag}_prof_fast -f $B --no-logmap --no-f64 > $o/${tag}_ncu_full.log 2>&1
tail -1 $o/${tag}_ncu_full.log | cut -c1-160
ncu --set full --clock-control none --import-source on -k regex:fast_s16 -s 7 -c 1 -o $o/${tag}_prof_logmap -f $B --no-f64 > $o/${tag}_ncu_logmap.log 2>&1
tail -1 $o/${tag}_ncu_logmap.log | cut -c1-160
for n in 512 4096 16384; do timeout 120 python tools/time_ref64.py $n 2>&1 | tail -1 | tee -a $o/${tag}_ref64_timing.txt; done
ncu --set full --clock-control none --import-source on -k regex:ref64_decode -s 2 -c 1 -o $o/${tag}_prof_ref64 -f python tools/time_ref64.py 4096 > $o/${tag}_ncu_ref64.log 2>&1
tail -1 $o/${tag}_ncu_ref64.log | cut -c1-160
