"""GPU debugging aid for the tcgen05 conv kernel: tiny GEMMs with structured inputs, prints what came back."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daclip_b200 import lib as L, ops

torch.manual_seed(0)
def run(tokens, cin, cout, kind):
    if kind == "ident":
        x = torch.zeros(1, 1, tokens, cin, device="cuda")
        x[0, 0, :, :] = torch.arange(tokens, device="cuda")[:, None] * 0 + torch.arange(cin, device="cuda")[None, :] * 0.01
        x[0, 0, :, 0] = torch.arange(tokens, device="cuda") * 1.0
        w = torch.eye(cout, cin, device="cuda")
    else:
        x = torch.randn(1, 1, tokens, cin, device="cuda")
        w = torch.randn(cout, cin, device="cuda") * cin ** -0.5
    xb = x.to(torch.bfloat16)
    out = torch.full((1, 1, tokens, cout), -777.0, device="cuda", dtype=torch.bfloat16)
    plan = ops.ConvPlan(xb, cin, ops.pack_linear(w), out, B=1, H=1, W=tokens)
    print("plan", plan.info(), flush=True)
    plan.run()
    torch.cuda.synchronize()
    ref = torch.nn.functional.linear(xb.float(), w.to(torch.bfloat16).float())
    o = out.float()
    print(kind, tokens, cin, cout, "untouched:", (o == -777).sum().item(), "nan:", torch.isnan(o).sum().item(),
          "maxerr:", (o - ref).abs().max().item(), flush=True)
    print(" out[0:4,0:8]\n", o[0, 0, :4, :8], "\n ref\n", ref[0, 0, :4, :8], flush=True)
    if tokens > 40:
        print(" out[33:35,0:8]\n", o[0, 0, 33:35, :8], "\n ref\n", ref[0, 0, 33:35, :8], flush=True)

for args in [(128, 64, 64, "ident"), (128, 64, 64, "rand"), (256, 128, 128, "rand"), (1000, 64, 256, "rand")]:
    try:
        run(*args)
    except Exception as e:
        print("EXC", args, e, flush=True)
        break
