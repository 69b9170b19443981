"""Runs a few representative conv plans repeatedly (for ncu / quick timing).  usage: prof_conv.py [case ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daclip_b200 import lib as L, ops

CASES = {
    # name: (B, H, W, cin, cout, kind)
    "l0_3x3": (16, 256, 256, 64, 64, "3x3"),
    "l0_3x3_cat": (16, 256, 256, 128, 64, "3x3"),
    "l0_qkv": (16, 256, 256, 64, 384, "qkv"),
    "l1_3x3": (16, 128, 128, 128, 128, "3x3"),
    "l2_3x3": (16, 64, 64, 256, 256, "3x3"),
    "l3_3x3": (16, 32, 32, 512, 512, "3x3"),
    "l3_geglu": (16, 32, 32, 512, 4096, "geglu"),
    "l0_1x1": (16, 256, 256, 128, 64, "1x1"),
    "l1_up": (16, 128, 128, 128, 64, "up"),
    "l0_down": (16, 256, 256, 64, 64, "down"),
    "l1_pair64": (16, 128, 128, 64, 64, "pair"),
    "l2_3x3_128": (16, 64, 64, 128, 128, "3x3"),
    "l3_3x3_256": (16, 32, 32, 256, 256, "3x3"),
    "l3_1x1_512": (16, 32, 32, 512, 512, "1x1"),
    "stem": (16, 256, 256, 64, 64, "stem"),
    "stem_pair": (16, 256, 128, 64, 128, "stem"),
    "l0_q": (16, 256, 256, 64, 128, "q"),
    "l0_kv": (16, 256, 256, 64, 256, "kv"),
    "l0_toout": (16, 256, 256, 128, 64, "toout"),
    "l0_qout": (16, 256, 256, 64, 64, "qout"),
    "l0_kvtc": (16, 256, 256, 64, 256, "kvtc"),
    "l0_qout_pn": (16, 256, 256, 64, 64, "qout_pn"),
    "l1_qout": (16, 128, 128, 64, 64, "qout"),
    "l1_qout_pn": (16, 128, 128, 64, 64, "qout_pn"),
    "l0_kvtc_pn": (16, 256, 256, 64, 256, "kvtc_pn"),
    "l1_kvtc_pn": (16, 128, 128, 64, 256, "kvtc_pn"),
    "l1_kvtc": (16, 128, 128, 64, 256, "kvtc"),
    "l0_pair": (16, 256, 256, 64, 64, "pair"),
    "l0_pair_cat": (16, 256, 256, 128, 64, "pair"),
    "l1_pair": (16, 128, 128, 64, 64, "pair"),
    "l0_pair_plain": (16, 256, 256, 64, 64, "pair_plain"),
    "l0_final_pair": (16, 256, 256, 64, 3, "final_pair"),
    "l0_final": (16, 256, 256, 64, 3, "final"),
    "l0_pair_res": (16, 256, 256, 64, 64, "pair_res"),
    "l0_pair_skip": (16, 256, 256, 64, 64, "pair_skip"),
    "l0_3x3_skip": (16, 256, 256, 64, 64, "3x3_skip"),
    "l1_3x3_64": (16, 128, 128, 64, 64, "3x3"),
    "l1_qout": (16, 128, 128, 128, 128, "qout"),
}

def make(name):
    B, H, W, cin, cout, kind = CASES[name]
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, H, W, cin, device="cuda", generator=g).to(torch.bfloat16)
    out_c = cout // 2 if kind == "geglu" else cout
    out = torch.zeros(B, H, W, out_c, device="cuda", dtype=torch.bfloat16)
    if kind == "3x3":
        w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * (9 * cin) ** -0.5
        film = torch.randn(B, 2 * cout, device="cuda", generator=g) * 0.1
        plan = ops.ConvPlan(x, cin, ops.pack_conv(w), out, B=B, H=H, W=W, act=L.ACT_SILU, film=film)
    elif kind == "qkv":
        w = torch.randn(cout, cin, device="cuda", generator=g) * cin ** -0.5
        q = torch.zeros(B, H, W, 128, device="cuda", dtype=torch.bfloat16)
        kv = torch.zeros(B, 256, H, W, device="cuda", dtype=torch.bfloat16)
        plan = ops.ConvPlan(x, cin, ops.pack_linear(w), q, B=B, H=H, W=W, epi=L.EPI_QKV, block_n=128, out_planar=kv)
    elif kind == "q":
        w = torch.randn(cout, cin, device="cuda", generator=g) * cin ** -0.5
        q = torch.zeros(B, H, W, 128, device="cuda", dtype=torch.bfloat16)
        plan = ops.ConvPlan(x, cin, ops.pack_linear(w), q, B=B, H=H, W=W, epi=L.EPI_QKV, block_n=128)
    elif kind == "kv":
        w = torch.randn(cout, cin, device="cuda", generator=g) * cin ** -0.5
        shift = torch.full((128,), 12.0, device="cuda")
        ctx = torch.zeros(B, 4, ops.ctx_slots(B, H, W, False), 32 * 34, device="cuda")
        plan = ops.ConvPlan(x, cin, ops.pack_linear(w), None, B=B, H=H, W=W, epi=L.EPI_KVCTX, block_n=256,
                            kv_shift=shift, ctx_acc=ctx)
    elif kind in ("pair_plain", "pair_res"):
        w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * (9 * cin) ** -0.5
        res = torch.randn(B, H, W, 64, device="cuda", generator=g).to(torch.bfloat16) if kind == "pair_res" else None
        plan = ops.PairConvPlan(x, ops.pack_conv_pair(w), out, B=B, H=H, W=W, act=L.ACT_SILU if res is not None else L.ACT_NONE,
                                res=res)
    elif kind in ("final", "final_pair"):
        w = torch.randn(3, cin, 3, 3, device="cuda", generator=g) * (9 * cin) ** -0.5
        b = torch.randn(3, device="cuda", generator=g)
        o32 = torch.zeros(B, 3, H, W, device="cuda")
        if kind == "final":
            plan = ops.ConvPlan(x, cin, ops.pack_conv(w), None, B=B, H=H, W=W, bias=b, out_nchw=o32)
        else:
            plan = ops.PairConvPlan(x, ops.pack_conv_pair(torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, 13))), None,
                                    B=B, H=H, W=W, bias=b, out_nchw=o32)
    elif kind == "pair":
        w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * (9 * cin) ** -0.5
        film = torch.randn(B, 2 * cout, device="cuda", generator=g) * 0.1
        a = x[..., :64].contiguous()
        s1 = x[..., 64:].contiguous() if cin == 128 else None
        plan = ops.PairConvPlan(a, ops.pack_conv_pair(w), out, B=B, H=H, W=W, src1=s1, act=L.ACT_SILU, film=film)
    elif kind in ("pair_skip", "3x3_skip"):
        w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * (9 * cin) ** -0.5
        wr = torch.randn(64, 128, device="cuda", generator=g) * 128 ** -0.5
        r0 = torch.randn(B, H, W, 64, device="cuda", generator=g).to(torch.bfloat16)
        r1 = torch.randn(B, H, W, 64, device="cuda", generator=g).to(torch.bfloat16)
        if kind == "pair_skip":
            plan = ops.PairConvPlan(x, ops.pack_conv_pair(w), out, B=B, H=H, W=W, act=L.ACT_SILU, rsrc0=r0, rsrc1=r1,
                                    rweight=ops.pack_linear(wr))
        else:
            plan = ops.ConvPlan(x, cin, ops.pack_conv(w), out, B=B, H=H, W=W, act=L.ACT_SILU, rsrc0=r0, rc0=64, rsrc1=r1,
                                rc1=64, rweight=ops.pack_linear(wr))
    elif kind in ("kvtc", "kvtc_pn"):
        w = torch.randn(cout, cin, device="cuda", generator=g) * cin ** -0.5
        shift = torch.full((128,), 12.0, device="cuda")
        pn = kind == "kvtc_pn"
        ctx = torch.zeros(B, 4, ops.ctx_slots(B, H, W, True), ops.KV_G_REC if pn else 32 * 34, device="cuda")
        plan = ops.KvPlan(x.reshape(B * H * W, cin), ops.centre_rows(w[:128]).to(torch.bfloat16).contiguous() if pn else ops.pack_kv_grouped(w),
                          shift, ctx, B, H * W, cin, prenorm_eps=1e-5 if pn else None)
        plan.info = lambda: {}
    elif kind == "toout":
        w = torch.randn(cout, cin, device="cuda", generator=g) * cin ** -0.5
        weff = (torch.randn(B, cout, cin, device="cuda", generator=g) * cin ** -0.5).to(torch.bfloat16)
        res = torch.randn(B, H, W, cout, device="cuda", generator=g).to(torch.bfloat16)
        plan = ops.ConvPlan(x, cin, ops.pack_linear(w), out, B=B, H=H, W=W, epi=L.EPI_LN,
                            bias=torch.zeros(cout, device="cuda"), ln_g=torch.ones(cout, device="cuda"), res=res,
                            per_image_w=True, weight_override=weff)
    elif kind == "qout_pn":
        wq = (torch.randn(1, 128, cin, device="cuda", generator=g) * cin ** -0.5).to(torch.bfloat16)
        weff = (torch.randn(B, cout, 128, device="cuda", generator=g) * 128 ** -0.5).to(torch.bfloat16)
        plan = ops.QoutPlan(x, wq, weff, x, out, torch.zeros(cout, device="cuda"), torch.ones(cout, device="cuda"),
                            1e-5, B, H * W, cout, prenorm_eps=1e-5, q_shift=torch.full((128,), 12.0, device="cuda"))
        plan.info = lambda: {}
    elif kind == "qout":
        wq = (torch.randn(1, 128, cin, device="cuda", generator=g) * cin ** -0.5).to(torch.bfloat16)
        weff = (torch.randn(B, cout, 128, device="cuda", generator=g) * 128 ** -0.5).to(torch.bfloat16)
        res = torch.randn(B, H, W, cout, device="cuda", generator=g).to(torch.bfloat16)
        plan = ops.QoutPlan(x, wq, weff, res, out, torch.zeros(cout, device="cuda"), torch.ones(cout, device="cuda"),
                            1e-5, B, H * W, cout)
        plan.info = lambda: {}
    elif kind == "up":
        w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * (9 * cin) ** -0.5
        out = torch.zeros(B, 2 * H, 2 * W, cout, device="cuda", dtype=torch.bfloat16)
        plan = ops.ConvPlan(x, cin, ops.pack_upsample_conv(w), out, B=B, H=H, W=W, bias=torch.zeros(cout, device="cuda"))
    elif kind == "down":
        w = torch.randn(cout, cin, 4, 4, device="cuda", generator=g) * (16 * cin) ** -0.5
        out = torch.zeros(B, H // 2, W // 2, cout, device="cuda", dtype=torch.bfloat16)
        plan = ops.ConvPlan(x, cin, ops.pack_conv(w, stride=2, pad=1), out, B=B, H=H, W=W, bias=torch.zeros(cout, device="cuda"))
    elif kind == "stem":
        w6 = torch.randn(64, 6, 7, 7, device="cuda", generator=g) * 0.05
        plan = ops.ConvPlan(x, 64, ops.pack_stem_pair(w6) if cout == 128 else ops.pack_stem(w6), out, B=B, H=H, W=W)
    elif kind == "geglu":
        w = torch.randn(cout, cin, device="cuda", generator=g) * cin ** -0.5
        b = torch.randn(cout, device="cuda", generator=g)
        pw, bp = ops.pack_geglu(w, b)
        plan = ops.ConvPlan(x, cin, pw, out, B=B, H=H, W=W, epi=L.EPI_GEGLU, bias=bp, block_n=256)
    else:
        w = torch.randn(cout, cin, device="cuda", generator=g) * cin ** -0.5
        plan = ops.ConvPlan(x, cin, ops.pack_linear(w), out, B=B, H=H, W=W)
    return plan

names = sys.argv[1:] or list(CASES)
for name in names:
    plan = make(name)
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
    print(f"{name:12s} {ms*1e3:9.1f} us  {plan.flops/ms/1e9:8.1f} TFLOP/s  {plan.info()}", flush=True)
