"""Aggregates warp-stall samples of an `ncu --page source --csv` dump into instruction windows (role-level view)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
k, s = ci["# Samples"], ci["Source"]
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
win = int(sys.argv[2]) if len(sys.argv) > 2 else 60
tot = sum(float(r[k] or 0) for r in data)
print("total", tot, "instructions", len(data))
for a in range(0, len(data), win):
    chunk = data[a:a + win]
    n = sum(float(r[k] or 0) for r in chunk)
    if n < 0.01 * tot:
        continue
    agg = {}
    for r in chunk:
        for j in stalls:
            agg[hdr[j]] = agg.get(hdr[j], 0) + float(r[j] or 0)
    top = sorted(agg.items(), key=lambda x: -x[1])[:2]
    hot = max(chunk, key=lambda r: float(r[k] or 0))
    tags = set()
    for r in chunk:
        for t in ("UTMALDG", "UTCHMMA", "LDTM", "STG", "MUFU", "TRYWAIT", "LDG"):
            if t in r[s]:
                tags.add(t)
    print(f"[{a:5d}] {100*n/tot:5.1f}%  {top[0][0]}:{100*top[0][1]/max(n,1):.0f}% {top[1][0]}:{100*top[1][1]/max(n,1):.0f}%  "
          f"tags={','.join(sorted(tags))}  hot={int(float(hot[k]))}:{hot[s].strip()[:60]}")
