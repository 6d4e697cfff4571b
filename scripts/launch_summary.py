"""Aggregates an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <cmd>`) by kernel
name: launches, total and mean device time, share of the listed time.  usage: launch_summary.py X.csv [skip_first_n_launches]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = [r for r in csv.reader(l for l in open(path, newline="") if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1 + skip:]:
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name)
    name = name if len(name) < 90 else name[:87] + "..."
    ns = float(r[vi].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ui], 1.0)
    agg[name][0] += 1
    agg[name][1] += ns
tot = sum(v[1] for v in agg.values())
print(f"{len(rows) - 1 - skip} launches, {tot / 1e6:.3f} ms of kernel time (cold-cache, serialised under ncu)")
print(f"{'kernel':90s} {'n':>6s} {'total us':>10s} {'mean us':>9s} {'share':>6s}")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:90s} {n:6d} {ns / 1e3:10.1f} {ns / 1e3 / n:9.1f} {100 * ns / tot:5.1f}%")
