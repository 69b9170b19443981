import os, sys
sys.path.insert(0, "/root/repo")
os.environ["DAC_CTA2"] = sys.argv[1]
os.environ["DAC_CTA2_DEBUG"] = "1"
import torch
from daclip_b200 import lib as L, ops
x = torch.randn(16, 32, 32, 512, device="cuda").to(torch.bfloat16)
w = torch.randn(512, 512, device="cuda") * 0.04
out = torch.zeros(16, 32, 32, 512, device="cuda", dtype=torch.bfloat16)
plan = ops.ConvPlan(x, 512, ops.pack_linear(w), out, B=16, H=32, W=32)
print("info", plan.info(), flush=True)
plan.run(); torch.cuda.synchronize()
ref = (x.float().reshape(-1, 512) @ w.to(torch.bfloat16).float().t()).reshape(out.shape)
print("cta2", sys.argv[1], "max err", (out.float() - ref).abs().max().item(), "ref max", ref.abs().max().item())
