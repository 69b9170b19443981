import os, sys, torch
sys.path.insert(0, "/root/repo")
import torch.nn.functional as F
from daclip_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
for (B, n, heads, grid) in [(1, 256, 2, 1), (1, 256, 2, 2), (1, 512, 4, 3), (1,4096,16,148)]:
    os.environ["DAC_ATTN_GRID"] = str(grid)
    qkv = torch.randn(B, n, 3 * heads * 32, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.zeros(B, n, heads * 32, device="cuda", dtype=torch.bfloat16)
    ops.attention(qkv, out, B, n, heads, 32)
    torch.cuda.synchronize()
    q, k, v = [t.reshape(B, n, heads, 32).transpose(1, 2) for t in qkv.float().chunk(3, dim=-1)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, n, heads * 32)
    print(B, n, heads, grid, "max err", (out.float() - ref).abs().max().item(), flush=True)
