"""Times dac_attention (d = 64) at the DA-CLIP ViT shapes: (B, tokens, heads) = (256, 50, 12) ViT-B/32, (64, 257, 16) ViT-L/14."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daclip_b200 import ops

g = torch.Generator(device="cuda").manual_seed(0)
shapes = [(256, 50, 12), (64, 257, 16)] if len(sys.argv) < 2 else [tuple(int(v) for v in sys.argv[1].split(","))]
for B, n, heads in shapes:
    qkv = torch.randn(B, n, 3 * heads * 64, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.zeros(B, n, heads * 64, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.attention(qkv, out, B, n, heads, 64)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ops.attention(qkv, out, B, n, heads, 64)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    byts = 2.0 * B * n * heads * 64 * 4
    print(f"B={B} n={n} heads={heads}: {ms*1e3:8.1f} us  {byts/ms/1e6:7.1f} GB/s algorithmic", flush=True)
