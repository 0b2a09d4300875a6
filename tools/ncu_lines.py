"""Per-instruction stall samples of an .ncu-rep: python tools/ncu_lines.py rep [min_pct] -- prints SASS lines whose
sample share is at least min_pct (default 0.3) with a running region total (regions split at backward branches)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp] or 0) for r in data)
reg_s, reg_n, reg_start = 0, 0, 0
for k, r in enumerate(data):
    s = int(r[isamp] or 0)
    reg_s += s; reg_n += 1
    if 100.0 * s / tot >= thr:
        print("%5d %5.2f%% x%-10s %s" % (k, 100.0 * s / tot, r[iex], r[isrc][:80]))
    if "BRA" in r[isrc] and reg_s * 100.0 / tot >= 1.0:
        print("   ---- region [%d,%d]: %.1f%% of samples, %d instructions, executed x%s" % (reg_start, k, 100.0 * reg_s / tot, reg_n, r[iex]))
    if "BRA" in r[isrc]:
        reg_s, reg_n, reg_start = 0, 0, k + 1
