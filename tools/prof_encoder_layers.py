"""Per-launch CUDA-event times of one DaCLIP.encode_image(control=True) (eager replay of the engine's steps)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daclip_b200 import synthetic
from daclip_b200.daclip import DaCLIP
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
clip = DaCLIP().load_reference_state_dict(synthetic.daclip_visual_state_dict(10)).to(dev).eval()
img = torch.randn(B, 3, 224, 224, device=dev)
for _ in range(2):
    clip.encode_image(img, control=True)
eng = next(iter(clip._engines.values()))
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(eng.steps) + 1)]
for rep in range(2):
    ev[0].record()
    for i, fn in enumerate(eng.steps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(len(eng.steps))]
# the steps of one block repeat: patchify | per tower: conv1, embed, 12 x (ln, qkv, attn, out, ln, fc, proj[, zero]), pool
print("steps", len(ts), "total us", sum(ts))
fused = os.environ.get("DAC_FUSE_ZERO", "1") != "0"   # zero-linears inside the CLIP tower's c_proj GEMMs: no "zero" step
per = 7 if fused else 8
names = ["ln1", "qkv", "attn", "out", "ln2", "fc", "proj", "zero"][:per]
agg = collections.defaultdict(float)
base = 3
for l in range(12):
    for j, n in enumerate(names):
        agg["ctl." + n] += ts[base + l * per + j]
base2 = base + 12 * per + 1 + 2
for l in range(12):
    for j, n in enumerate(names[:7]):
        agg["clip." + n] += ts[base2 + l * 7 + j]
for k, v in agg.items():
    print(f"{k:10s} {v:9.1f} us  ({v/12:7.1f} per layer)")
print("first steps", [round(t, 1) for t in ts[:4]], "pool", round(ts[base + 12 * per], 1))
