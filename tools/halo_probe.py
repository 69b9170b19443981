"""Experiment: 3x3 conv with ONE haloed activation load per K chunk (nine shifted UMMA operand views, row-group stride
(tile_w+2)*128 B).  Prints the max error against F.conv2d for both descriptor conventions and the timing at level-0
size.  usage: halo_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from daclip_b200 import lib as L, ops

torch.backends.cudnn.allow_tf32 = False
g = torch.Generator(device="cuda").manual_seed(0)
for (B, H, W, cin, cout) in [(2, 32, 32, 64, 64), (1, 40, 24, 128, 64)]:
    x = torch.randn(B, cin, H, W, device="cuda", generator=g)
    w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * (9 * cin) ** -0.5
    xh = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    ref = F.conv2d(xh.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), padding=1)
    for halo in (0, 1, 2):
        out = torch.zeros(B, H, W, cout, device="cuda", dtype=torch.bfloat16)
        plan = ops.ConvPlan(xh, cin, ops.pack_conv(w), out, B=B, H=H, W=W, halo=halo)
        plan.run()
        torch.cuda.synchronize()
        err = (out.float().permute(0, 3, 1, 2) - ref).abs().max().item()
        print(f"{B}x{H}x{W} {cin}->{cout} halo={halo}: max err {err:.4g}  {plan.info()}", flush=True)

B, H, W = 16, 256, 256
for cin, cout in [(64, 64), (128, 64)]:
    x = torch.randn(B, H, W, cin, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * (9 * cin) ** -0.5
    film = torch.randn(B, 2 * cout, device="cuda", generator=g) * 0.1
    out = torch.zeros(B, H, W, cout, device="cuda", dtype=torch.bfloat16)
    for halo in (0, 1, 2):
        plan = ops.ConvPlan(x, cin, ops.pack_conv(w), out, B=B, H=H, W=W, act=L.ACT_SILU, film=film, halo=halo)
        for _ in range(3):
            plan.run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            plan.run()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"l0 {cin}->{cout} halo={halo}: {ms*1e3:8.1f} us {plan.flops/ms/1e9:7.1f} TFLOP/s {plan.info()}", flush=True)
