"""times the captured step graph of whatever daclip_b200 is in cwd"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from daclip_b200 import synthetic
from daclip_b200.unet import ConditionalUNet
B, H, W = 16, 256, 256
sd, kw = synthetic.unet_state_dict(0)
inp = {k: v.cuda() for k, v in synthetic.restoration_inputs(B, H, W, T=1, seed=3).items()}
net = ConditionalUNet(**kw); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
eng = net.engine(B, H, W)
eng.set_inputs(inp["lq"], inp["lq"], inp["text_context"], inp["image_context"]); eng.set_time(37.0)
eng.replay(); torch.cuda.synchronize()
for _ in range(30): eng.replay()
res = []
for rnd in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(60): eng.replay()
    e1.record(); torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 60)
print(os.path.basename(os.getcwd()), " ".join(f"{r:.4f}" for r in res), "ms", flush=True)
