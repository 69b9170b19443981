"""Prints the hottest SASS instructions (warp-stall samples) of an `ncu --page source --csv` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
k = ci["# Samples"]
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    try:
        data.append((float(r[k] or 0), r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for s, r in data:
    for i in stalls:
        agg[hdr[i]] = agg.get(hdr[i], 0) + float(r[i] or 0)
print("stall mix:", sorted(((round(100 * v / max(tot, 1), 1), n) for n, v in agg.items() if v), reverse=True)[:8])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
order = sorted(range(len(data)), key=lambda i: -data[i][0])[:n]
for i in sorted(order):
    s, r = data[i]
    top = sorted(((float(r[j] or 0), hdr[j]) for j in stalls), reverse=True)[0]
    print(f"{i:5d} {int(s):6d} {100*s/tot:5.1f}%  {r[ci['Source']].strip()[:90]:90s} {top[1]}")
