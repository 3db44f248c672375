"""Per-source-line cost of a kernel in an .ncu-rep (needs -lineinfo + --import-source on; no GPU needed).
usage: python tools/ncu_lines.py rep [min_pct]   -> lines with >= min_pct of the warp instructions, grouped by file"""
import csv, io, subprocess, sys

rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
files, cur, hdr = [], None, None
for row in csv.reader(io.StringIO(out)):
    if not row:
        continue
    if row[0] in ("File Name", "File Path"):
        cur = {"name": row[1], "rows": []}
        files.append(cur)
        hdr = None
    elif row[0] == "Line No":
        hdr = row
    elif cur is not None and hdr is not None and row[0].isdigit() and row[2] in ("", "-"):  # source line rows have no SASS address
        cur["rows"].append(row)
iI, iT, iN = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
def num(x):
    try: return int(x)
    except ValueError: return 0
tot = sum(num(r[iI]) for f in files for r in f["rows"]); tots = sum(num(r[iN]) for f in files for r in f["rows"])
print(f"total {tot} warp instructions, {tots} samples")
for f in files:
    ft = sum(num(r[iI]) for r in f["rows"]); fs = sum(num(r[iN]) for r in f["rows"])
    if ft == 0: continue
    print(f"== {f['name']}: {100*ft/tot:.1f}% of instructions, {100*fs/max(tots,1):.1f}% of samples")
    for r in f["rows"]:
        n = num(r[iI])
        if 100 * n / tot >= min_pct or 100 * num(r[iN]) / max(tots, 1) >= min_pct:
            print(f"  {r[0]:>5s} {100*n/tot:5.2f}% thr {num(r[iT])/max(n,1):5.1f} smp {100*num(r[iN])/max(tots,1):5.2f}%  {r[1].strip()[:130]}")
