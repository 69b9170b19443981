"""A/B timing of the UNet step graph under engine switches on ONE box (clocks differ between boxes):
usage: ab_engine.py [B H W] [SWITCH] -- times the captured evaluation with UNetEngine.SWITCH (default PDL) off and on,
interleaved, and reports how far the outputs of the two settings are apart."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daclip_b200 import synthetic
from daclip_b200.unet import ConditionalUNet, UNetEngine

B, H, W = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (16, 256, 256)
SWITCH = sys.argv[4] if len(sys.argv) >= 5 else "PDL"
sd, kw = synthetic.unet_state_dict(0)
inp = {k: v.cuda() for k, v in synthetic.restoration_inputs(B, H, W, T=1, seed=3).items()}


def build(**switches):
    for k, v in switches.items():
        setattr(UNetEngine, k, v)
    net = ConditionalUNet(**kw)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    eng = net.engine(B, H, W)
    eng.set_inputs(inp["lq"], inp["lq"], inp["text_context"], inp["image_context"])
    eng.set_time(37.0)
    eng.replay()
    torch.cuda.synchronize()
    return net, eng


def timeit(eng, n=40):
    for _ in range(5):
        eng.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        eng.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


variants = {"pdl_off": {SWITCH: False}, "pdl_off2": {SWITCH: False}, "pdl_on": {SWITCH: True}}
engines = {k: build(**v) for k, v in variants.items()}
ref = engines["pdl_off"][1].out_noise.clone()
for k, (_, e) in engines.items():
    print(k, "max |diff| vs pdl_off", (e.out_noise - ref).abs().max().item())
e = engines["pdl_on"][1]
for i in range(3):
    e.replay()
    torch.cuda.synchronize()
    print("pdl_on replay", i, "max |diff| vs pdl_off", (e.out_noise - ref).abs().max().item(), "out max", ref.abs().max().item())
for rnd in range(3):
    print(" ".join(f"{k} {timeit(e):.4f} ms" for k, (_, e) in engines.items()), flush=True)
