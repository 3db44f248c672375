"""Summarise an .ncu-rep here (no GPU): key raw metrics per kernel + SASS opcode histogram + stall totals.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-index]"""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")][:90])
    for w in WANT:
        if w in hdr:
            print(f"   {w:85s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the source page concatenates kernels; split on the "Kernel Name" rows
blocks, cur = [], None
for row in csv.reader(io.StringIO(src)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
b = blocks[min(which, len(blocks) - 1)]
h = b["hdr"]
iS, iI, iT, iN = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
ops = collections.Counter(); thr = collections.Counter(); smp = collections.Counter()
tot = 0
for r in b["rows"]:
    parts = r[iS].split()
    op = parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "?")
    op = op.rstrip(";")
    n = int(r[iI]); ops[op] += n; thr[op] += int(r[iT]); smp[op] += int(r[iN]); tot += n
print(f"== SASS opcode histogram of kernel #{which} ({b['name'][:60]}): {tot} warp instructions, {len(b['rows'])} SASS lines")
for op, n in ops.most_common(28):
    print(f"   {op:28s} {n:12d} {100*n/tot:5.1f}%  avg threads {thr[op]/max(n,1):5.1f}  samples {smp[op]}")
stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
tots = {c: sum(int(r[h.index(c)]) for r in b["rows"]) for c in stall_cols}
s = sum(tots.values())
print("== stall samples:", ", ".join(f"{c[6:]} {100*v/s:.1f}%" for c, v in sorted(tots.items(), key=lambda kv: -kv[1]) if v * 50 > s))
