"""Back-to-back timing of the small helper kernels at SpatialTransformer shapes (16384 tokens x 512)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daclip_b200 import ops

def timeit(name, fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    print(f"{name:34s} {a.elapsed_time(b) / n * 1e3:8.1f} us", flush=True)

B, hw = 16, 1024
for C in (512, 256):
    x = torch.randn(B, hw, C, device="cuda").to(torch.bfloat16)
    y = torch.empty_like(x)
    g = torch.ones(C, device="cuda"); bb = torch.zeros(C, device="cuda")
    stats = torch.zeros(B * 64, device="cuda")
    timeit(f"layernorm_rows gain-only C={C}", lambda: ops.layernorm_rows(x, y, B * hw, C, g, None, 1e-5))
    timeit(f"layernorm_rows affine C={C}", lambda: ops.layernorm_rows(x, y, B * hw, C, g, bb, 1e-5))
    timeit(f"groupnorm C={C}", lambda: ops.groupnorm_nhwc(x, y, B, hw, C, g, bb, stats))
x = torch.randn(16, 65536, 64, device="cuda").to(torch.bfloat16); y = torch.empty_like(x)
timeit("layernorm_rows L0 prenorm C=64", lambda: ops.layernorm_rows(x, y, 16 * 65536, 64, None, None, 1e-5))
x = torch.randn(16, 16384, 128, device="cuda").to(torch.bfloat16); y = torch.empty_like(x)
timeit("layernorm_rows L1 prenorm C=128", lambda: ops.layernorm_rows(x, y, 16 * 16384, 128, None, None, 1e-5))
