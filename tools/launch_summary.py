"""Condense an ncu launch list (`ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv
--log-file X ...`) into one row per kernel: launches, total time, share of the listed time, DRAM GB.
usage: launch_summary.py launches.csv summary.csv"""
import csv
import re
import sys
from collections import OrderedDict

src, out = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
per = OrderedDict()      # launch id -> [name, time_ns, bytes]
for r in rows[1:]:
    e = per.setdefault(r[ix["ID"]], [r[ix["Kernel Name"]], 0.0, 0.0])
    v = float(r[ix["Metric Value"]].replace(",", ""))
    if r[ix["Metric Name"]] == "gpu__time_duration.sum":
        e[1] += v
    elif r[ix["Metric Name"]].startswith("dram__bytes"):
        e[2] += v


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.split("::")[-1]


agg = OrderedDict()
for name, t, b in per.values():
    a = agg.setdefault(short(name), [0, 0.0, 0.0])
    a[0] += 1; a[1] += t; a[2] += b
total = sum(a[1] for a in agg.values())
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "total_time_us", "share_pct", "dram_GB"])
    for k, (n, t, b) in agg.items():
        w.writerow([k, n, f"{t / 1e3:.1f}", f"{100 * t / total:.1f}", f"{b / 1e9:.3f}"])
    w.writerow(["TOTAL", sum(a[0] for a in agg.values()), f"{total / 1e3:.1f}", "100.0", f"{sum(a[2] for a in agg.values()) / 1e9:.3f}"])
print(open(out).read())
