"""Times dac_attention (d = 32) at the SpatialTransformer shapes.  DAC_NO_TC_ATTN=1 selects the mma.sync kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daclip_b200 import ops

g = torch.Generator(device="cuda").manual_seed(0)
for B, n, heads in [(16, 1024, 16), (16, 1024, 8), (8, 4096, 16)]:
    qkv = torch.randn(B, n, 3 * heads * 32, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.zeros(B, n, heads * 32, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.attention(qkv, out, B, n, heads, 32)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ops.attention(qkv, out, B, n, heads, 32)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    fl = 4.0 * B * heads * n * n * 32
    ex = B * heads * n * n
    print(f"B={B} n={n} heads={heads}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s  {ex/ms/1e6/148:6.2f} exp/ns/SM "
          f"(SFU peak 16/clk)", flush=True)
