"""Summarise an .ncu-rep (read here, no GPU needed): python tools/ncu_summary.py rep [--src]"""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for i, h in enumerate(hdr):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
        print(f"{h:88s} {vals[i]:>16s} {units[i]}")
if "--src" in sys.argv:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr, data = rows[1], rows[2:]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    tot_e = sum(int(r[iex] or 0) for r in data)
    tot_s = sum(int(r[isamp] or 0) for r in data)
    print("SASS instructions:", len(data), " warp-instructions executed:", tot_e)
    ops = collections.Counter()
    for r in data:
        op = r[isrc].split()[1] if r[isrc].startswith("@") else r[isrc].split()[0]
        ops[op] += int(r[iex] or 0)
    print("opcode mix (% of executed):", [(k, round(100 * v / tot_e, 1)) for k, v in ops.most_common(18)])
    B = 400
    for b in range(0, len(data), B):
        ch = data[b:b + B]
        s = sum(int(r[isamp] or 0) for r in ch); e = sum(int(r[iex] or 0) for r in ch)
        if e * 200 > tot_e or s * 200 > tot_s:
            print(f"  sass[{b:5d}:{b+B:5d}] samples {100*s/tot_s:5.1f}%  executed {100*e/tot_e:5.1f}%")
