python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/sweep_subblock.py 5824 5888 6016 6080 4032 2016 1632 704 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['K'], 'auto_L', r['auto_L'], 'auto', r['auto'])
"
