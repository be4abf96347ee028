"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share, average."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
    name = re.sub(r"\(.*", "", r[ki])
    agg[name][0] += 1
    agg[name][1] += float(r[vi].replace(",", "")) * scale
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':62s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:62]:62s} {v[0]:8d} {v[1] / 1e3:10.3f} {v[1] / tot * 100:6.1f}% {v[1] / v[0]:9.1f}")
print(f"{'TOTAL':62s} {sum(v[0] for v in agg.values()):8d} {tot / 1e3:10.3f}")
