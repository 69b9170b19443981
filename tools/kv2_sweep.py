"""Fixed cost vs per-tile cost of the in-kernel-PreNorm LinearAttention kernels: time over a sweep of sizes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from daclip_b200 import ops
C = 64
g = torch.Generator(device="cuda").manual_seed(0)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for B, H, W in [(1, 16, 16), (16, 16, 16), (16, 32, 32), (16, 64, 64), (16, 128, 128), (16, 256, 256), (8, 512, 512)]:
    x = torch.randn(B, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(256, C, device="cuda", generator=g) * C ** -0.5
    ns = ops.ctx_slots(B, H, W, True)
    ctx = torch.zeros(B, 4, ns, ops.KV_G_REC, device="cuda")
    plan = ops.KvPlan(x.reshape(-1, C), ops.centre_rows(w[:128]).to(torch.bfloat16).contiguous(), torch.full((128,), 12.0, device="cuda"),
                      ctx, B, H * W, C, prenorm_eps=1e-5)
    t_kv = timeit(plan.run)
    t_ms = timeit(lambda: ctx.zero_())
    weff = torch.randn(B, 64, 128, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.empty_like(x)
    pq = ops.QoutPlan(x, ops.pack_linear(ops.centre_rows(w[:128]).contiguous()).w, weff, x, out, torch.zeros(C, device="cuda"),
                      torch.ones(C, device="cuda"), 1e-5, B, H * W, C, prenorm_eps=1e-5, q_shift=torch.full((128,), 12.0, device="cuda"))
    t_q = timeit(pq.run)
    tiles = B * H * W // 128
    print(f"B={B} {H}x{W}: tiles {tiles:6d} ({tiles/148:6.1f}/CTA) slots {ns:3d} ctx {ctx.numel()*4/1e6:6.2f} MB  kv {t_kv:7.1f} us  memset {t_ms:6.1f} us  qout {t_q:7.1f} us", flush=True)
