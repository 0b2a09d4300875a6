"""Per-region (runs of SASS lines with equal execution count) sample / instruction shares of an ncu source page CSV."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def I(r, k):
    v = r[ix[k]]
    return int(v) if v else 0
S = sum(I(r, '# Samples') for r in data); E = sum(I(r, 'Instructions Executed') for r in data)
seg = []; start = 0
for n in range(1, len(data) + 1):
    if n == len(data) or I(data[n], 'Instructions Executed') != I(data[start], 'Instructions Executed'):
        seg.append((start, n)); start = n
keys = ['stall_selected', 'stall_wait', 'stall_math', 'stall_not_selected', 'stall_branch_resolving', 'stall_no_inst',
        'stall_dispatch', 'stall_short_sb', 'stall_long_sb', 'stall_barrier']
for a, b in seg:
    s = sum(I(r, '# Samples') for r in data[a:b]); e = sum(I(r, 'Instructions Executed') for r in data[a:b])
    if s / S > thr:
        st = {k: sum(I(r, k) for r in data[a:b]) for k in keys}
        tt = sum(st.values()) or 1
        print(f"[{a:5d},{b:5d}) n={b-a:4d} exec/inst={I(data[a],'Instructions Executed')//1000:4d}k samp={100*s/S:5.1f}% exec={100*e/E:5.1f}% rel.cost={s/S/(e/E) if e else 0:4.2f}",
              {k[6:]: round(100 * v / tt) for k, v in st.items() if v / tt > 0.04})
