#!/bin/bash
# full GPU check: every -m gpu test, then one short bench line
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_quick.json').readlines()[-1])
print('BENCH value', d['value'], 'e2e', d['e2e']['value'], 'int8', d['e2e']['with_int8_llrs']['value'], 'alu', d['roofline']['alu'])
print('LM', d.get('logmap_s16'))
print('ET', d.get('early_termination'))
print('CPU', d.get('cpu_baseline'))
"
