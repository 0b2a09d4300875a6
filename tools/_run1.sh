python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
r = json.loads(sys.stdin.read())
print(json.dumps({'value': r['value'], 'e2e': r['e2e'], 'frac_alu': r['roofline']['alu']['frac']}))"
