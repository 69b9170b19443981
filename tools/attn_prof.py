import os, sys, torch
sys.path.insert(0, "/root/repo")
from daclip_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
B, n, heads = 16, 1024, 16
qkv = torch.randn(B, n, 3 * heads * 32, device="cuda", generator=g).to(torch.bfloat16)
out = torch.zeros(B, n, heads * 32, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, out, B, n, heads, 32)
torch.cuda.synchronize()
os.environ["DAC_ATTN2_PROF_DUMP"] = "1"
ops.attention(qkv, out, B, n, heads, 32)
torch.cuda.synchronize()
