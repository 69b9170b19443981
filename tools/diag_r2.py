"""Round-2 diagnostic: where does the epilogue time go?  Runs conv / attention cases with the profiling-only debug masks
(DAC_EPI_DEBUG: 1 no stores, 2 no activation, 4 no FiLM; DAC_ATTN_DEBUG: 1 no exponentials, 2 no max pass, 4 no P store)
and prints one table.  Results with a mask != 0 are WRONG by design - timing only."""
import os, sys, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0], "__none__"]
import torch
from daclip_b200 import ops

def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

# prof_conv.make without running its main loop
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "prof_conv.py")).read().split("names = sys.argv")[0]
ns = {"__file__": os.path.join(os.path.dirname(os.path.abspath(__file__)), "prof_conv.py")}
exec(compile(src, "prof_conv_defs", "exec"), ns)
make = ns["make"]

print("== conv epilogue decomposition (us) ==")
cases = ["l0_pair", "l0_pair_res", "l0_pair_cat", "l0_pair_skip", "l0_pair_plain", "l0_3x3", "l1_3x3", "l3_3x3", "l3_geglu", "l0_1x1"]
masks = [0, 1, 2, 4, 7]
print(f"{'case':14s}" + "".join(f"  dbg={m:<3d}" for m in masks))
for c in cases:
    row = []
    for m in masks:
        os.environ["DAC_EPI_DEBUG"] = str(m)
        plan = make(c)
        row.append(timeit(plan.run))
        del plan
    print(f"{c:14s}" + "".join(f"  {t:7.1f}" for t in row), flush=True)
os.environ["DAC_EPI_DEBUG"] = "0"

print("== attention decomposition (us) ==")
g = torch.Generator(device="cuda").manual_seed(0)
for B, n, heads in [(16, 1024, 16), (16, 1024, 8), (8, 4096, 16)]:
    qkv = torch.randn(B, n, 3 * heads * 32, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.zeros(B, n, heads * 32, device="cuda", dtype=torch.bfloat16)
    row = []
    for m in [0, 1, 2, 4, 3, 7]:
        os.environ["DAC_ATTN_DEBUG"] = str(m)
        row.append(timeit(lambda: ops.attention(qkv, out, B, n, heads, 32)))
    print(f"B={B} n={n} h={heads}: " + "  ".join(f"dbg{m}={t:7.1f}" for m, t in zip([0, 1, 2, 4, 3, 7], row)), flush=True)
os.environ["DAC_ATTN_DEBUG"] = "0"
