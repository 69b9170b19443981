"""Replays the UNet step graph several times on the same inputs and reports, in execution order, which block outputs
(engine.taps) are not bit-identical between replays.  usage: determinism.py [B H W] [SWITCH=0/1 ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daclip_b200 import synthetic
from daclip_b200.unet import ConditionalUNet, UNetEngine

args = [a for a in sys.argv[1:] if "=" not in a]
for a in sys.argv[1:]:
    if "=" in a:
        k, v = a.split("=")
        setattr(UNetEngine, k, bool(int(v)))
        print("switch", k, bool(int(v)))
B, H, W = (int(a) for a in args[:3]) if len(args) >= 3 else (16, 256, 256)
sd, kw = synthetic.unet_state_dict(0)
inp = {k: v.cuda() for k, v in synthetic.restoration_inputs(B, H, W, T=1, seed=3).items()}
net = ConditionalUNet(**kw)
net.load_state_dict(sd, strict=True)
net = net.cuda().eval()
eng = net.engine(B, H, W)
eng.set_inputs(inp["lq"], inp["lq"], inp["text_context"], inp["image_context"])
eng.set_time(37.0)
snaps = []
for i in range(4):
    eng.replay()
    torch.cuda.synchronize()
    snaps.append({k: v.clone() for k, v in eng.taps.items()} | {"out_noise": eng.out_noise.clone()})
bad = 0
for k in snaps[0]:
    d = max((snaps[i][k].float() - snaps[0][k].float()).abs().max().item() for i in range(1, 4))
    if d != 0.0:
        bad += 1
        print(f"{k:32s} max |replay_i - replay_0| = {d:.3e}   (|x| max {snaps[0][k].float().abs().max().item():.3e})")
print("non-deterministic taps:", bad, "of", len(snaps[0]))
