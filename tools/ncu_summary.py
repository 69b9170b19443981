"""Key metrics + per-region stall histogram of an .ncu-rep (run where ncu is installed).  usage: ncu_summary.py rep [window]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
W = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum"]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:60])
    for k in keys:
        if k in hdr:
            print(f"   {k:85s} {units[hdr.index(k)]:12s} {r[hdr.index(k)]}")
    for i, k in enumerate(hdr):
        if "shared" in k and "tc" in k and k not in keys:
            print(f"   {k:85s} {units[i]:12s} {r[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
k, ex = ci["# Samples"], ci["Instructions Executed"]
tot = sum(float(r[k] or 0) for r in data)
print("samples", tot, "instructions", len(data))
for s in range(0, len(data), W):
    blk = data[s:s + W]
    sm = sum(float(r[k] or 0) for r in blk)
    if sm < 0.004 * tot:
        continue
    exs = max(float(r[ex] or 0) for r in blk)
    ops = {}
    for r in blk:
        t = r[ci["Source"]].strip().split()
        if not t:
            continue
        o = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        o = o.split(".")[0]
        ops[o] = ops.get(o, 0) + 1
    top = sorted(ops.items(), key=lambda x: -x[1])[:4]
    hot = max(blk, key=lambda r: float(r[k] or 0))
    print(f"{s:5d} {sm:7.0f} {100*sm/tot:5.1f}%  maxexec {exs:9.0f}  {top}   hot: {hot[ci['Source']].strip()[:60]}")
