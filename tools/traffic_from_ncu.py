"""profiles/traffic.json from an `ncu --set full` report: DRAM bytes (read + write) of the longest captured launch of the
dominant kernel.  usage: python tools/traffic_from_ncu.py <kind: bic|cbic> <report.ncu-rep> <kernel regex> [summary.txt]"""
import csv
import json
import os
import re
import subprocess
import sys

kind, rep, pat = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(h)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def val(r, name):
    return float(r[col[name]].replace(",", "")) * scale.get(units[col[name]], 1.0)


launches = []
for r in rows[2:]:
    if not re.search(pat, r[col["Kernel Name"]]):
        continue
    launches.append(dict(kernel=re.sub(r"\(.*", "", r[col["Kernel Name"]]), us=val(r, "gpu__time_duration.sum"),
                         read=val(r, "dram__bytes_read.sum"), write=val(r, "dram__bytes_write.sum"), grid=int(float(r[col["launch__grid_size"]])),
                         occ=float(r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]])))
lines = [f"{'kernel':44s} {'grid':>8s} {'us':>9s} {'dram rd MB':>11s} {'dram wr MB':>11s} {'GB/s':>8s} {'warps act %':>11s}"]
for l in launches:
    lines.append(f"{l['kernel'][:44]:44s} {l['grid']:8d} {l['us']:9.1f} {l['read'] / 1e6:11.1f} {l['write'] / 1e6:11.1f} "
                 f"{(l['read'] + l['write']) / l['us'] / 1e3:8.0f} {l['occ']:11.1f}")
print("\n".join(lines))
if len(sys.argv) > 4:
    open(sys.argv[4], "w").write("\n".join(lines) + "\n")
top = max(launches, key=lambda l: l["us"])
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[kind] = {"dram_bytes_per_launch": top["read"] + top["write"], "kernel": top["kernel"], "launch_us_under_ncu": top["us"], "grid": top["grid"],
              "dram_gbps_under_ncu": (top["read"] + top["write"]) / top["us"] / 1e3, "report": os.path.basename(rep),
              "note": "longest captured launch of the dominant kernel; ncu timings are cold-cache and serialised"}
json.dump(data, open(path, "w"), indent=1)
