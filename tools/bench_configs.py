"""Throughput of the other BASELINE.json configs (C1 B=1 256^2, C4 B=8 512^2, C5 encoder B=256) - report numbers for
DESIGN.md; the contract benchmark is bench.py."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daclip_b200 import synthetic
from daclip_b200.daclip import DaCLIP
from daclip_b200.sde import IRSDE
from daclip_b200.unet import ConditionalUNet

dev = torch.device("cuda:0")
sd, kw = synthetic.unet_state_dict(0)
net = ConditionalUNet(**kw); net.load_state_dict(sd); net = net.to(dev).eval()
sde = IRSDE(50, T=100, schedule="cosine", eps=0.005, device=dev); sde.set_model(net)
out = {}
for name, B, S in [("C1_b1_256", 1, 256), ("C4_b8_512", 8, 512), ("C2_b16_256", 16, 256)]:
    inp = {k: v.to(dev) for k, v in synthetic.restoration_inputs(B, S, S, T=1, seed=5).items()}
    sde.set_mu(inp["lq"])
    x_T = inp["lq"] + inp["eps0"] * sde.max_sigma
    kwargs = dict(text_context=inp["text_context"], image_context=inp["image_context"])
    for _ in range(2):
        sde.reverse_posterior(x_T, **kwargs)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); n = 2
    for _ in range(n):
        sde.reverse_posterior(x_T, **kwargs)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    out[name] = {"images_per_s": round(B / (ms / 1e3), 3), "ms_per_denoiser_step": round(ms / 100, 3)}
    net.invalidate() if False else None
clip = DaCLIP().load_reference_state_dict(synthetic.daclip_visual_state_dict(10)).to(dev).eval()
img = torch.randn(256, 3, 224, 224, device=dev)
for _ in range(3):
    clip.encode_image(img, control=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    clip.encode_image(img, control=True)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
out["C5_encoder_b256"] = {"ms_per_batch": round(ms, 3), "images_per_s": round(256 / (ms / 1e3), 1),
                          "tflops": round(256 * 18.34e9 / (ms / 1e3) / 1e12, 1)}
print(json.dumps(out))
