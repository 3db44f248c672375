"""List the hottest contiguous SASS regions of a kernel in an .ncu-rep (no GPU needed).
usage: python tools/ncu_hot.py rep [kernel-index] [min_exec_fraction]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
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
tot = sum(int(r[iI]) for r in b["rows"])
tots = sum(int(r[iN]) for r in b["rows"])
print(f"kernel {b['name'][:80]}: {tot} warp instr, {tots} samples")
for k, r in enumerate(b["rows"]):
    n = int(r[iI])
    print(f"{k:5d} {100*n/tot:5.2f}% thr {int(r[iT])/max(n,1):5.1f} smp {100*int(r[iN])/tots:5.2f}%  {r[iS].strip()}")
