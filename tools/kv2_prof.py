"""Per-warp wait totals of linattn_kv2_kernel (needs the DAC_KV2_PROF variant: DAC_LIB=da-clip_b200/libdac_b200_kvprof.so)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from daclip_b200 import ops
B, H, W, C = 16, 256, 256, 64
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
w = torch.randn(256, C, device="cuda", generator=g) * C ** -0.5
ctx = torch.zeros(B, 4, ops.ctx_slots(B, H, W, True), ops.KV_G_REC, device="cuda")
plan = ops.KvPlan(x.reshape(-1, C), ops.centre_rows(w[:128]).to(torch.bfloat16).contiguous(), torch.full((128,), 12.0, device="cuda"),
                  ctx, B, H * W, C, prenorm_eps=1e-5)
for _ in range(3):
    plan.run()
torch.cuda.synchronize()
os.environ["DAC_KV2_PROF_DUMP"] = "1"
print("roles: 0 producer [empty]; 1 issuer [full acc_empty stat p_full g_flushed]; 2-5 stats [full]; 6-21 epilogue [acc_full tmem_ld stat p_free exp stage fence+arrive]; last = total")
plan.run()
torch.cuda.synchronize()
