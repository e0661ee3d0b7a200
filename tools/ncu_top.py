"""Top stall locations of one kernel from an ncu source-page CSV:
   ncu -i rep.ncu-rep --page source --csv --launch-skip N --launch-count 1 > src.csv ; python tools/ncu_top.py src.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print(rows[0][:2])
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
iS, iSrc, iEx = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


tot = sum(num(r[iS]) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for r in data:
    for h in stalls:
        agg[h] = agg.get(h, 0) + num(r[hdr.index(h)])
print("stall mix:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for r in sorted(data, key=lambda r: -num(r[iS]))[:n]:
    st = {h: num(r[hdr.index(h)]) for h in stalls}
    main = sorted(((v, k) for k, v in st.items() if v), reverse=True)[:3]
    print(f"{num(r[iS]):6d} {100 * num(r[iS]) / max(tot, 1):5.1f}% ex={r[iEx]:>9s} {r[iSrc][:78]:78s} {main}")
