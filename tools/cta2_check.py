"""CTA-pair mode (DAC_CTA2=1) against the 1-CTA kernel on the same inputs: outputs must agree bit for bit (same MMA
order per accumulator element), plus timings.  usage: cta2_check.py [case ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daclip_b200 import lib as L, ops

g = torch.Generator(device="cuda").manual_seed(0)


def rnd(*shape, scale=1.0):
    return torch.randn(*shape, device="cuda", generator=g) * scale


def make(case, cta2):
    os.environ["DAC_CTA2"] = "1" if cta2 else "0"
    os.environ["DAC_CTA2_RES"] = "1" if cta2 else "0"
    kind, B, H, W, cin, cout = case
    x = CACHE.setdefault(("x", case), rnd(B, H, W, cin).to(torch.bfloat16))
    if kind == "1x1":
        w = CACHE.setdefault(("w", case), rnd(cout, cin, scale=cin ** -0.5))
        b = CACHE.setdefault(("b", case), rnd(cout))
        out = torch.zeros(B, H, W, cout, device="cuda", dtype=torch.bfloat16)
        plan = ops.ConvPlan(x, cin, ops.pack_linear(w), out, B=B, H=H, W=W, bias=b, act=L.ACT_GELU)
    elif kind == "3x3":
        w = CACHE.setdefault(("w", case), rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5))
        film = CACHE.setdefault(("f", case), rnd(B, 2 * cout, scale=0.1))
        out = torch.zeros(B, H, W, cout, device="cuda", dtype=torch.bfloat16)
        plan = ops.ConvPlan(x, cin, ops.pack_conv(w), out, B=B, H=H, W=W, act=L.ACT_SILU, film=film)
    elif kind == "3x3res":
        w = CACHE.setdefault(("w", case), rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5))
        res = CACHE.setdefault(("r", case), rnd(B, H, W, cout).to(torch.bfloat16))
        out = torch.zeros(B, H, W, cout, device="cuda", dtype=torch.bfloat16)
        plan = ops.ConvPlan(x, cin, ops.pack_conv(w), out, B=B, H=H, W=W, act=L.ACT_SILU, res=res)
    elif kind in ("pair", "pair_res", "pair_cat"):
        w = CACHE.setdefault(("w", case), rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5))
        film = CACHE.setdefault(("f", case), rnd(B, 2 * cout, scale=0.1))
        res = CACHE.setdefault(("r", case), rnd(B, H, W, cout).to(torch.bfloat16))
        out = torch.zeros(B, H, W, cout, device="cuda", dtype=torch.bfloat16)
        a = x[..., :64].contiguous()
        s1 = x[..., 64:].contiguous() if cin == 128 else None
        if kind == "pair_res":
            plan = ops.PairConvPlan(a, ops.pack_conv_pair(w), out, B=B, H=H, W=W, act=L.ACT_SILU, res=res)
        else:
            plan = ops.PairConvPlan(a, ops.pack_conv_pair(w), out, B=B, H=H, W=W, src1=s1, act=L.ACT_SILU, film=film)
    elif kind in ("pair_skip", "skip"):
        w = CACHE.setdefault(("w", case), rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5))
        rc = 128 if kind == "pair_skip" else cout + cout // 2
        wr = CACHE.setdefault(("wr", case), rnd(cout, rc, scale=rc ** -0.5))
        r0 = CACHE.setdefault(("r0", case), rnd(B, H, W, cout).to(torch.bfloat16))
        r1 = CACHE.setdefault(("r1", case), rnd(B, H, W, rc - cout).to(torch.bfloat16))
        out = torch.zeros(B, H, W, cout, device="cuda", dtype=torch.bfloat16)
        if kind == "pair_skip":
            plan = ops.PairConvPlan(x, ops.pack_conv_pair(w), out, B=B, H=H, W=W, act=L.ACT_SILU, rsrc0=r0, rsrc1=r1,
                                    rweight=ops.pack_linear(wr))
        else:
            plan = ops.ConvPlan(x, cin, ops.pack_conv(w), out, B=B, H=H, W=W, act=L.ACT_SILU, rsrc0=r0, rc0=cout, rsrc1=r1,
                                rc1=rc - cout, rweight=ops.pack_linear(wr))
    elif kind == "up":
        w = CACHE.setdefault(("w", case), rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5))
        b = CACHE.setdefault(("b", case), rnd(cout))
        out = torch.zeros(B, 2 * H, 2 * W, cout, device="cuda", dtype=torch.bfloat16)
        plan = ops.ConvPlan(x, cin, ops.pack_upsample_conv(w), out, B=B, H=H, W=W, bias=b)
    elif kind == "down":
        w = CACHE.setdefault(("w", case), rnd(cout, cin, 4, 4, scale=(16 * cin) ** -0.5))
        b = CACHE.setdefault(("b", case), rnd(cout))
        out = torch.zeros(B, H // 2, W // 2, cout, device="cuda", dtype=torch.bfloat16)
        plan = ops.ConvPlan(x, cin, ops.pack_conv(w, stride=2, pad=1), out, B=B, H=H, W=W, bias=b)
    elif kind == "halo":
        w = CACHE.setdefault(("w", case), rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5))
        film = CACHE.setdefault(("f", case), rnd(B, 2 * cout, scale=0.1))
        out = torch.zeros(B, H, W, cout, device="cuda", dtype=torch.bfloat16)
        os.environ["DAC_NO_PAIR"] = "1"
        plan = ops.ConvPlan(x, cin, ops.pack_conv(w), out, B=B, H=H, W=W, act=L.ACT_SILU, film=film)
    elif kind == "geglu":
        w = CACHE.setdefault(("w", case), rnd(cout, cin, scale=cin ** -0.5))
        b = CACHE.setdefault(("b", case), rnd(cout))
        pw, bp = ops.pack_geglu(w, b)
        out = torch.zeros(B, H, W, cout // 2, device="cuda", dtype=torch.bfloat16)
        plan = ops.ConvPlan(x, cin, pw, out, B=B, H=H, W=W, epi=L.EPI_GEGLU, bias=bp, block_n=256)
    elif kind == "f32":
        w = CACHE.setdefault(("w", case), rnd(cout, cin, scale=cin ** -0.5))
        b = CACHE.setdefault(("b", case), rnd(cout))
        stream = CACHE.setdefault(("s", case), rnd(B, H, W, cout)).clone()
        out = stream
        plan = ops.ConvPlan(x, cin, ops.pack_linear(w), None, B=B, H=H, W=W, bias=b, res_f32=stream, out_f32=stream)
    plan._out = out
    return plan


CACHE = {}
CASES = {
    "lin512": ("1x1", 16, 32, 32, 512, 512),
    "vit_qkv": ("1x1", 1, 1, 12800, 768, 2304),
    "vit_fc": ("1x1", 1, 1, 12800, 768, 3072),
    "vit_out": ("f32", 1, 1, 12800, 768, 768),
    "vit_proj": ("f32", 1, 1, 12800, 3072, 768),
    "l3_256": ("3x3", 16, 32, 32, 256, 256),
    "l2_128": ("3x3", 16, 64, 64, 128, 128),
    "l3_512": ("3x3res", 16, 32, 32, 512, 512),
    "l2_256": ("3x3", 16, 64, 64, 256, 256),
    "geglu": ("geglu", 16, 32, 32, 512, 4096),
    "odd": ("3x3", 2, 40, 24, 128, 128),
    "p64": ("pair", 16, 256, 256, 64, 64),
    "p64res": ("pair_res", 16, 256, 256, 64, 64),
    "p128": ("pair_cat", 16, 256, 256, 128, 64),
    "p64l1": ("pair", 16, 128, 128, 64, 64),
    "p64small": ("pair", 2, 40, 24, 64, 64),
    "p64skip": ("pair_skip", 16, 256, 256, 64, 64),
    "up128": ("up", 16, 128, 128, 128, 64),
    "up512": ("up", 16, 32, 32, 512, 256),
    "down64": ("down", 16, 256, 256, 64, 64),
    "down128": ("down", 16, 64, 64, 128, 256),
    "halo64": ("halo", 16, 128, 128, 64, 64),
    "upsmall": ("up", 2, 20, 12, 128, 64),
    "skip128": ("skip", 16, 128, 128, 128, 128),
}
for name in (sys.argv[1:] or list(CASES)):
    case = CASES[name]
    res = {}
    for cta2 in (0, 1):
        plan = make(case, cta2)
        plan.run()
        torch.cuda.synchronize()
        first = plan._out.clone()
        for _ in range(3):
            plan.run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            plan.run()
        b.record()
        torch.cuda.synchronize()
        res[cta2] = (first, a.elapsed_time(b) / 10 * 1e3, plan.info())
    d = (res[0][0].float() - res[1][0].float()).abs().max().item()
    print(f"{name:10s} 1-CTA {res[0][1]:7.1f} us  pair {res[1][1]:7.1f} us  max|diff| {d:.3g}  finite {bool(torch.isfinite(res[1][0].float()).all())}  "
          f"{res[0][2]} -> {res[1][2]}", flush=True)
