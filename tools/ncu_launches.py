"""Summarises an ncu launch list (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`)
per kernel: launches, mean / total duration, mean DRAM bytes - and writes the JSON bench.py reads its `roofline.traffic`
from.  usage: ncu_launches.py launches.csv out.json BATCH IMAGE"""
import csv, json, re, sys
src, dst, B, S = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
per = {}
for r in rows[1:]:
    if r[ci["ID"]] == "ID":
        continue
    name = re.sub(r"\(.*", "", r[ci["Kernel Name"]]).replace("dac::", "").replace("void ", "")
    name = re.sub(r"<.*", "", name)
    k = per.setdefault((r[ci["ID"]], name), {})
    val = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    m = r[ci["Metric Name"]]
    if m.startswith("gpu__time_duration"):
        val *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)       # -> us
    else:
        val *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    k[m] = val
agg = {}
for (_, name), m in per.items():
    a = agg.setdefault(name, {"launches": 0, "us": 0.0, "dram": 0.0})
    a["launches"] += 1
    a["us"] += m.get("gpu__time_duration.sum", 0.0)
    a["dram"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
tot = sum(a["us"] for a in agg.values())
out = {"batch": B, "image": S, "source": src, "total_us": round(tot, 1), "kernels": {}}
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    out["kernels"][name] = {"launches": a["launches"], "us_total": round(a["us"], 1), "share": round(a["us"] / tot, 4),
                            "us_per_launch": round(a["us"] / a["launches"], 2),
                            "dram_bytes_per_launch": round(a["dram"] / a["launches"], 1)}
    print(f"{a['us']/tot*100:5.1f} %  {a['launches']:5d} x {a['us']/a['launches']:8.1f} us  {a['dram']/a['launches']/1e6:8.1f} MB  {name}")
gem = [a for n, a in agg.items() if n.startswith(("conv_igemm_kernel", "linattn_kv", "linattn_qout"))]
if gem:
    n = sum(a["launches"] for a in gem)
    out["gemm_kernels"] = {"launches": n, "dram_bytes_per_launch": round(sum(a["dram"] for a in gem) / n, 1),
                           "share": round(sum(a["us"] for a in gem) / tot, 4)}
json.dump(out, open(dst, "w"), indent=1)
