"""Top stall locations of an ncu --import-source report:  ncu -i rep --page source --csv | python tools/ncu_top_stalls.py [N]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hdr]
ci = {n: i for i, n in enumerate(h)}
key = "# Samples"
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
data = []
for r in rows[hdr + 1:]:
    try:
        v = float(r[ci[key]])
    except Exception:
        continue
    data.append((v, r))
tot = sum(v for v, _ in data) or 1
agg = {n: 0.0 for n in stalls}
for v, r in data:
    for n in stalls:
        try:
            agg[n] += float(r[ci[n]])
        except Exception:
            pass
print("stall reasons:", ", ".join(f"{n[6:]}={x / sum(agg.values()) * 100:.0f}%" for n, x in sorted(agg.items(), key=lambda kv: -kv[1])[:6]))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 25
for v, r in sorted(data, key=lambda t: -t[0])[:N]:
    top = max(stalls, key=lambda n: float(r[ci[n]] or 0))
    print(f"{v / tot * 100:5.1f}%  {top[6:]:12s} {r[ci['Source']][:140]}")
