"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: count, total
device time, share.  python tools/launch_summary.py gpurun_out/launches.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki])
    v = float(r[vi].replace(",", ""))
    unit = r[ui]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':60s} {'launches':>9s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {n:9d} {us / 1e3:10.2f} {us / n:10.1f} {100 * us / tot:6.1f}%")
print(f"{'total':60s} {sum(v[0] for v in agg.values()):9d} {tot / 1e3:10.2f}")
