"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.  usage: python tools/launch_list.py x.csv"""
import collections, csv, sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[start:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    ms = v / 1e6 if r[ui] == "ns" else v / 1e3 if r[ui] in ("us", "usecond") else v
    a = agg.setdefault(r[ki][:78], [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
print(f"# cold-cache, serialised per-launch times: compare SHARES, not absolutes.  total {tot:.2f} ms over {sum(a[0] for a in agg.values())} launches")
print(f"{'kernel':80s} {'count':>5s} {'total ms':>10s} {'share':>7s} {'ms/launch':>9s}")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:80s} {n:5d} {ms:10.3f} {100*ms/tot:6.1f}% {ms/n:9.3f}")
