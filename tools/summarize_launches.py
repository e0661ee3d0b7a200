"""Summarise an ncu per-launch device-time list (ncu --metrics gpu__time_duration.sum --csv) by kernel."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
hdr = None
for r in csv.reader(open(sys.argv[1])):
    if r and r[0] == "ID":
        hdr = r
        break
iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in rows:
    v = float(r[iV].replace(",", ""))
    u = r[iU]
    us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    name = re.sub(r"\(.*", "", r[iK]).replace("void ", "").replace("pn::", "")
    agg[name][0] += 1
    agg[name][1] += us
    tot += us
print(f"{len(rows)} launches, {tot / 1e3:.2f} ms device time (serialised, cold-cache: compare SHARES)")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us / 1e3:9.3f} ms {100 * us / tot:5.1f}%  x{n:5d}  {name}")
