"""GPU parity tests, kernel by kernel, through the C ABI (daclip_b200.ops -> libdac_b200.so).

Checker: plain fp32 PyTorch ops (TF32 disabled) / the functional oracle in oracle/, on the same seeded
inputs.  Tolerances: SDE updates are bit-exact (fp32, reference operation order); bf16 tensor-core layers are
compared on bf16-rounded inputs with fp32 accumulation, so the only differences are accumulation order and the
final bf16 rounding of the output: |err| <= 2^-8 * max(1, |ref|) + small absolute slack.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def bf(x):
    return x.to(torch.bfloat16)


def nhwc(x):  # NCHW fp32 -> NHWC bf16
    return bf(x.permute(0, 2, 3, 1)).contiguous()


def nchw(x):  # NHWC bf16 -> NCHW fp32
    return x.float().permute(0, 3, 1, 2).contiguous()


def assert_close_bf16(got, ref, name, rel=2 ** -7, abs_=2e-3):
    got, ref = got.float(), ref.float()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    err = (got - ref).abs()
    tol = rel * ref.abs() + abs_ * max(1.0, ref.abs().max().item())
    bad = (err > tol).sum().item()
    assert bad == 0 and torch.isfinite(got).all(), (
        f"{name}: {bad}/{err.numel()} elements out of tolerance, max err {err.max().item():.4g} "
        f"(ref max {ref.abs().max().item():.4g})")


@pytest.fixture(scope="module")
def ops(cuda):
    from daclip_b200 import lib, ops as o
    lib.load()
    return o


def rnd(gen, *shape, scale=1.0):
    return (torch.randn(*shape, generator=gen, device="cuda") * scale)


@pytest.fixture()
def gen(cuda):
    g = torch.Generator(device="cuda")
    g.manual_seed(1234)
    return g


# ---------------------------------------------------------------------------------------------- conv / linear
@pytest.mark.parametrize("tokens,cin,cout", [(1000, 64, 64), (333, 128, 256), (4096, 512, 1536), (50 * 7, 768, 768)])
def test_linear_plain(ops, gen, tokens, cin, cout):
    from daclip_b200 import lib as L
    x = bf(rnd(gen, 1, 1, tokens, cin))
    w = rnd(gen, cout, cin, scale=cin ** -0.5)
    b = rnd(gen, cout)
    pw = ops.pack_linear(w)
    out = torch.full((1, 1, tokens, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    plan = ops.ConvPlan(x, cin, pw, out, B=1, H=1, W=tokens, bias=b)
    plan.run()
    torch.cuda.synchronize()
    ref = F.linear(x.float(), bf(w).float(), b)
    assert_close_bf16(out, ref, f"linear {tokens}x{cin}->{cout} {plan.info()}")


@pytest.mark.parametrize("B,H,W,cin,cout,tile", [
    (2, 32, 32, 64, 64, None), (1, 16, 48, 128, 128, None), (2, 64, 64, 64, 64, (8, 16)),
    (1, 256, 256, 64, 64, None), (2, 32, 32, 512, 512, None), (1, 6, 6, 256, 512, None)])
def test_conv3x3_film_silu(ops, gen, B, H, W, cin, cout, tile):
    from daclip_b200 import lib as L
    x = rnd(gen, B, cin, H, W)
    w = rnd(gen, cout, cin, 3, 3, scale=(9 * cin) ** -0.5)
    film = rnd(gen, B, 2 * cout + 8, scale=0.5)
    xh = nhwc(x)
    out = torch.full((B, H, W, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    plan = ops.ConvPlan(xh, cin, ops.pack_conv(w), out, B=B, H=H, W=W, act=L.ACT_SILU, film=film, film_off=8,
                        tile=tile)
    plan.run()
    torch.cuda.synchronize()
    y = F.conv2d(nchw(xh), bf(w).float(), padding=1)
    sc = film[:, 8:8 + cout, None, None]
    sh = film[:, 8 + cout:8 + 2 * cout, None, None]
    ref = F.silu(y * (sc + 1) + sh)
    assert_close_bf16(nchw(out), ref, f"conv3x3 {B}x{H}x{W} {cin}->{cout} {plan.info()}")


@pytest.mark.parametrize("B,H,W,concat,mode", [
    (2, 32, 32, False, "film"), (1, 16, 48, True, "film"), (3, 40, 24, False, "res"), (2, 37, 50, True, "plain"),
    (1, 256, 256, False, "film"), (5, 16, 16, False, "res"), (1, 8, 16, True, "res")])
def test_conv3x3_pixel_pair(ops, gen, B, H, W, concat, mode):
    """Pixel-pair mode (dac_conv_desc.pair): the tensors viewed as [B, H, W/2, 2C], centre taps as N = 128 MMAs against
    overlapping windows of one 192-row weight block per ky, side taps as N = 64 MMAs into one half of the accumulator;
    FiLM vectors shared by the two pixels of a pair; ragged sizes (partial tiles, odd pair counts)."""
    from daclip_b200 import lib as L
    cin = 128 if concat else 64
    x = rnd(gen, B, cin, H, W)
    w = rnd(gen, 64, cin, 3, 3, scale=(9 * cin) ** -0.5)
    xh = nhwc(x)
    a = xh[..., :64].contiguous()
    s_ = xh[..., 64:].contiguous() if concat else None
    film = rnd(gen, B, 2 * 64 + 8, scale=0.5) if mode == "film" else None
    res = nhwc(rnd(gen, B, 64, H, W)) if mode == "res" else None
    out = torch.full((B, H, W, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    act = L.ACT_NONE if mode == "plain" else L.ACT_SILU
    plan = ops.PairConvPlan(a, ops.pack_conv_pair(w), out, B=B, H=H, W=W, src1=s_, act=act, film=film, film_off=8,
                            res=res)
    for _ in range(2):
        plan.run()
    torch.cuda.synchronize()
    y = F.conv2d(nchw(xh), bf(w).float(), padding=1)
    if mode == "film":
        y = F.silu(y * (film[:, 8:72, None, None] + 1) + film[:, 72:136, None, None])
    elif mode == "res":
        y = F.silu(y) + nchw(res)
    assert_close_bf16(nchw(out), y, f"pixel-pair conv3x3 {B}x{H}x{W} {cin}->64 {mode} {plan.info()}")


@pytest.mark.parametrize("B,H,W,Hc,Wc", [(2, 32, 32, 32, 32), (1, 48, 64, 40, 57), (3, 16, 16, 16, 15), (1, 256, 256, 256, 256)])
def test_final_conv_pixel_pair_nchw(ops, gen, B, H, W, Hc, Wc):
    """final_conv (64 -> 3, fp32 NCHW output cropped to the un-padded image) in pixel-pair mode: N = 32 centre MMAs over
    16-row weight blocks, two 16-column halves in the epilogue."""
    x = nhwc(rnd(gen, B, 64, H, W))
    w = rnd(gen, 3, 64, 3, 3, scale=(9 * 64) ** -0.5)
    b = rnd(gen, 3)
    out = torch.full((B, 3, Hc, Wc), float("nan"), device="cuda")
    wp = ops.pack_conv_pair(torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, 13)))
    plan = ops.PairConvPlan(x, wp, None, B=B, H=H, W=W, bias=b, out_nchw=out)
    for _ in range(2):
        plan.run()
    torch.cuda.synchronize()
    ref = F.conv2d(nchw(x), bf(w).float(), b, padding=1)[:, :, :Hc, :Wc]
    assert (out - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item()), (out - ref).abs().max().item()


@pytest.mark.parametrize("B,H,W,two", [(2, 32, 32, True), (3, 24, 40, False), (1, 256, 256, True), (2, 19, 34, True)])
def test_conv3x3_pixel_pair_fused_skip(ops, gen, B, H, W, two):
    """Pixel-pair mode with the ResBlock's 1x1 res_conv fused as a second accumulator (even pixels -> first half of its
    columns, odd pixels -> second half): silu(conv3x3(h)) + W_r [x | skip]."""
    from daclip_b200 import lib as L
    rc = 128 if two else 64
    hh = nhwc(rnd(gen, B, 64, H, W))
    xr = nhwc(rnd(gen, B, rc, H, W))
    w = rnd(gen, 64, 64, 3, 3, scale=(9 * 64) ** -0.5)
    wr = rnd(gen, 64, rc, scale=rc ** -0.5)
    out = torch.full((B, H, W, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    r0 = xr[..., :64].contiguous()
    r1 = xr[..., 64:].contiguous() if two else None
    plan = ops.PairConvPlan(hh, ops.pack_conv_pair(w), out, B=B, H=H, W=W, act=L.ACT_SILU, rsrc0=r0, rsrc1=r1,
                            rweight=ops.pack_linear(wr))
    for _ in range(2):
        plan.run()
    torch.cuda.synchronize()
    ref = F.silu(F.conv2d(nchw(hh), bf(w).float(), padding=1)) + F.conv2d(nchw(xr), bf(wr).float()[:, :, None, None])
    assert_close_bf16(nchw(out), ref, f"pixel-pair conv3x3 + fused skip {B}x{H}x{W} rc={rc} {plan.info()}")


@pytest.mark.parametrize("B,H,W,c0,c1,cout", [(2, 40, 24, 64, 0, 64), (3, 37, 51, 64, 64, 64), (1, 16, 8, 128, 0, 64),
                                               (2, 9, 200, 64, 0, 128)])
def test_conv3x3_halo_load(ops, gen, B, H, W, c0, c1, cout):
    """ONE haloed (16+2) x (8+2) activation load per K chunk, nine shifted operand views (1280 B row-group stride):
    same result as the three-column-load path up to the fp32 accumulation order of the taps, on ragged sizes."""
    a = nhwc(rnd(gen, B, c0, H, W))
    s_ = nhwc(rnd(gen, B, c1, H, W)) if c1 else None
    w = rnd(gen, cout, c0 + c1, 3, 3, scale=(9 * (c0 + c1)) ** -0.5)
    outs = []
    for halo in (1, 0):
        out = torch.full((B, H, W, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
        plan = ops.ConvPlan(a, c0, ops.pack_conv(w), out, B=B, H=H, W=W, src1=s_, c1=c1, halo=halo,
                            tile=None if halo else (16, 8))
        plan.run()
        torch.cuda.synchronize()
        outs.append(out)
    x = nchw(a) if s_ is None else torch.cat([nchw(a), nchw(s_)], 1)
    assert_close_bf16(nchw(outs[0]), F.conv2d(x, bf(w).float(), padding=1), f"conv3x3 halo {plan.info()}")
    assert_close_bf16(outs[0], outs[1], "halo vs column loads", rel=2 ** -7, abs_=1e-3)


@pytest.mark.parametrize("share", [True, False])
@pytest.mark.parametrize("B,H,W,c0,c1,cout", [(2, 40, 24, 64, 0, 64), (1, 64, 64, 128, 64, 128), (1, 16, 16, 256, 0, 256)])
def test_conv3x3_tap_sharing_modes(ops, gen, share, B, H, W, c0, c1, cout):
    """One activation load per column of taps (row-shifted descriptors, resident weights when they fit) must give
    the same result as one load per tap."""
    a = nhwc(rnd(gen, B, c0, H, W))
    s_ = nhwc(rnd(gen, B, c1, H, W)) if c1 else None
    w = rnd(gen, cout, c0 + c1, 3, 3, scale=(9 * (c0 + c1)) ** -0.5)
    out = torch.full((B, H, W, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    plan = ops.ConvPlan(a, c0, ops.pack_conv(w), out, B=B, H=H, W=W, src1=s_, c1=c1, share_taps=share)
    plan.run()
    torch.cuda.synchronize()
    x = nchw(a) if s_ is None else torch.cat([nchw(a), nchw(s_)], 1)
    ref = F.conv2d(x, bf(w).float(), padding=1)
    assert_close_bf16(nchw(out), ref, f"conv3x3 share={share} {plan.info()}")


def test_conv3x3_concat_residual(ops, gen):
    from daclip_b200 import lib as L
    B, H, W, c0, c1, cout = 3, 16, 16, 64, 128, 128
    a, s = rnd(gen, B, c0, H, W), rnd(gen, B, c1, H, W)
    w = rnd(gen, cout, c0 + c1, 3, 3, scale=(9 * (c0 + c1)) ** -0.5)
    r = rnd(gen, B, cout, H, W)
    ah, sh_, rh = nhwc(a), nhwc(s), nhwc(r)
    out = torch.zeros(B, H, W, cout, device="cuda", dtype=torch.bfloat16)
    plan = ops.ConvPlan(ah, c0, ops.pack_conv(w), out, B=B, H=H, W=W, src1=sh_, c1=c1, act=L.ACT_SILU, res=rh)
    plan.run()
    torch.cuda.synchronize()
    ref = F.silu(F.conv2d(torch.cat([nchw(ah), nchw(sh_)], 1), bf(w).float(), padding=1)) + nchw(rh)
    assert_close_bf16(nchw(out), ref, "conv3x3 concat+res")


@pytest.mark.parametrize("B,H,W,c,rc0,rc1", [(2, 32, 32, 64, 64, 64), (1, 40, 24, 128, 128, 64), (1, 16, 16, 64, 64, 0)])
def test_conv3x3_fused_skip_conv(ops, gen, B, H, W, c, rc0, rc1):
    """ResBlock tail: out = SiLU(conv3x3(h)) + res_conv(cat[x, skip]) with the 1x1 product in its own TMEM columns."""
    from daclip_b200 import lib as L
    hh = nhwc(rnd(gen, B, c, H, W))
    x = nhwc(rnd(gen, B, rc0, H, W))
    sk = nhwc(rnd(gen, B, rc1, H, W)) if rc1 else None
    w = rnd(gen, c, c, 3, 3, scale=(9 * c) ** -0.5)
    wr = rnd(gen, c, rc0 + rc1, 1, 1, scale=(rc0 + rc1) ** -0.5)
    out = torch.full((B, H, W, c), float("nan"), device="cuda", dtype=torch.bfloat16)
    plan = ops.ConvPlan(hh, c, ops.pack_conv(w), out, B=B, H=H, W=W, act=L.ACT_SILU,
                        rsrc0=x, rc0=rc0, rsrc1=sk, rc1=rc1, rweight=ops.pack_linear(wr))
    plan.run()
    torch.cuda.synchronize()
    xin = nchw(x) if sk is None else torch.cat([nchw(x), nchw(sk)], 1)
    ref = F.silu(F.conv2d(nchw(hh), bf(w).float(), padding=1)) + F.conv2d(xin, bf(wr).float())
    assert_close_bf16(nchw(out), ref, f"fused skip conv {plan.info()}")


def test_conv_channel_slices(ops, gen):
    """Source read as a channel sub-range of a wider buffer; output written into a channel sub-range."""
    B, H, W = 2, 8, 16
    big = bf(rnd(gen, B, H, W, 384))
    w = rnd(gen, 64, 128, scale=128 ** -0.5)
    out = torch.zeros(B, H, W, 192, device="cuda", dtype=torch.bfloat16)
    src = big[..., 128:]                      # view: data_ptr offset, pixel pitch stays 384
    plan = ops.ConvPlan(src, 128, ops.pack_linear(w), out, B=B, H=H, W=W, ld0=384, out_coff=64)
    plan.run()
    torch.cuda.synchronize()
    ref = F.linear(big[..., 128:256].float(), bf(w).float())
    assert_close_bf16(out[..., 64:128], ref, "channel-slice linear")
    assert out[..., :64].abs().max() == 0 and out[..., 128:].abs().max() == 0


@pytest.mark.parametrize("B,H,W,cin,cout", [(2, 32, 32, 64, 128), (1, 64, 64, 128, 256), (1, 256, 256, 64, 64)])
def test_downsample_4x4_s2(ops, gen, B, H, W, cin, cout):
    x = rnd(gen, B, cin, H, W)
    w = rnd(gen, cout, cin, 4, 4, scale=(16 * cin) ** -0.5)
    b = rnd(gen, cout)
    xh = nhwc(x)
    out = torch.full((B, H // 2, W // 2, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    plan = ops.ConvPlan(xh, cin, ops.pack_conv(w, stride=2, pad=1), out, B=B, H=H, W=W, bias=b)
    plan.run()
    torch.cuda.synchronize()
    ref = F.conv2d(nchw(xh), bf(w).float(), b, stride=2, padding=1)
    assert_close_bf16(nchw(out), ref, f"downsample {H}x{W} {cin}->{cout}")


@pytest.mark.parametrize("B,H,W,cin,cout", [(2, 16, 16, 128, 64), (1, 32, 32, 512, 256), (1, 3, 5, 64, 64)])
def test_upsample_conv_folded(ops, gen, B, H, W, cin, cout):
    x = rnd(gen, B, cin, H, W)
    w = rnd(gen, cout, cin, 3, 3, scale=(9 * cin) ** -0.5)
    b = rnd(gen, cout)
    xh = nhwc(x)
    out = torch.full((B, 2 * H, 2 * W, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    plan = ops.ConvPlan(xh, cin, ops.pack_upsample_conv(w), out, B=B, H=H, W=W, bias=b)
    plan.run()
    torch.cuda.synchronize()
    ref = F.conv2d(F.interpolate(nchw(xh), scale_factor=2, mode="nearest"), bf(w).float(), b, padding=1)
    # folded taps are sums of up to 4 bf16-rounded-once weights: slightly looser than a plain conv
    assert_close_bf16(nchw(out), ref, f"upsample-conv {H}x{W} {cin}->{cout}", rel=2 ** -6, abs_=6e-3)


@pytest.mark.parametrize("B,H,W", [(2, 32, 32), (1, 24, 40), (1, 256, 256)])
def test_stem_and_init_conv(ops, gen, B, H, W):
    xt, cond = rnd(gen, B, 3, H, W), rnd(gen, B, 3, H, W)
    w = rnd(gen, 64, 6, 7, 7, scale=(49 * 6) ** -0.5)
    Hp, Wp = -(-H // 16) * 16, -(-W // 16) * 16
    stem = torch.full((B, Hp, Wp, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.stem_input(xt, cond, stem, H, W)
    out = torch.full((B, Hp, Wp, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    plan = ops.ConvPlan(stem, 64, ops.pack_stem(w), out, B=B, H=Hp, W=Wp)
    plan.run()
    torch.cuda.synchronize()
    x = torch.cat([xt - cond, cond], 1)
    x = F.pad(x, (0, Wp - W, 0, Hp - H), mode="reflect") if (Hp > H or Wp > W) else x
    ref = F.conv2d(bf(x).float(), bf(w).float(), padding=3)
    assert_close_bf16(nchw(out), ref, f"init_conv {H}x{W}")
    # pixel-pair form: one packed row per pair of adjacent pixels, 128 weight rows; the [B,Hp,Wp/2,128] output IS [B,Hp,Wp,64]
    stem2 = torch.full((B, Hp, Wp // 2, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.stem_input(xt, cond, stem2, H, W, pair=True)
    out2 = torch.full((B, Hp, Wp, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    plan2 = ops.ConvPlan(stem2, 64, ops.pack_stem_pair(w), out2.view(B, Hp, Wp // 2, 128), B=B, H=Hp, W=Wp // 2)
    plan2.run()
    torch.cuda.synchronize()
    assert_close_bf16(nchw(out2), ref, f"init_conv (pixel pairs) {H}x{W}")


def test_final_conv_nchw_crop(ops, gen):
    B, H, W, Hp, Wp = 2, 24, 40, 32, 48
    x = rnd(gen, B, 64, Hp, Wp)
    w, b = rnd(gen, 3, 64, 3, 3, scale=(9 * 64) ** -0.5), rnd(gen, 3)
    xh = nhwc(x)
    out = torch.full((B, 3, H, W), float("nan"), device="cuda", dtype=torch.float32)
    plan = ops.ConvPlan(xh, 64, ops.pack_conv(w), None, B=B, H=Hp, W=Wp, bias=b, out_nchw=out)
    plan.run()
    torch.cuda.synchronize()
    ref = F.conv2d(nchw(xh), bf(w).float(), b, padding=1)[..., :H, :W]
    assert (out - ref).abs().max().item() < 1e-3, (out - ref).abs().max().item()


def test_epilogue_ln_residual(ops, gen):
    from daclip_b200 import lib as L
    B, H, W, cin, cout = 2, 16, 32, 128, 128
    x, r = bf(rnd(gen, B, H, W, cin)), bf(rnd(gen, B, H, W, cout))
    w, b, g = rnd(gen, cout, cin, scale=cin ** -0.5), rnd(gen, cout), 1 + 0.1 * rnd(gen, cout)
    out = torch.zeros(B, H, W, cout, device="cuda", dtype=torch.bfloat16)
    plan = ops.ConvPlan(x, cin, ops.pack_linear(w), out, B=B, H=H, W=W, epi=L.EPI_LN, bias=b, ln_g=g, res=r)
    plan.run()
    torch.cuda.synchronize()
    y = F.linear(x.float(), bf(w).float(), b)
    ref = (y - y.mean(-1, keepdim=True)) * torch.rsqrt(y.var(-1, unbiased=False, keepdim=True) + 1e-5) * g + r.float()
    assert_close_bf16(out, ref, "LN epilogue")


def test_epilogue_geglu(ops, gen):
    from daclip_b200 import lib as L
    B, n, c = 2, 256, 256
    x = bf(rnd(gen, B, 1, n, c))
    w, b = rnd(gen, 8 * c, c, scale=c ** -0.5), rnd(gen, 8 * c)
    pw, bperm = ops.pack_geglu(w, b)
    out = torch.zeros(B, 1, n, 4 * c, device="cuda", dtype=torch.bfloat16)
    plan = ops.ConvPlan(x, c, pw, out, B=B, H=1, W=n, epi=L.EPI_GEGLU, bias=bperm, block_n=256)
    plan.run()
    torch.cuda.synchronize()
    y = F.linear(x.float(), bf(w).float(), b)
    val, gate = y.chunk(2, dim=-1)
    assert_close_bf16(out, val * F.gelu(gate), "GEGLU epilogue")


def test_epilogue_bias_img_two_residuals_gelu(ops, gen):
    from daclip_b200 import lib as L
    B, n, c = 3, 200, 256
    x, r1, r2 = bf(rnd(gen, B, 1, n, c)), bf(rnd(gen, B, 1, n, c)), bf(rnd(gen, B, 1, n, c))
    w, b, bi = rnd(gen, c, c, scale=c ** -0.5), rnd(gen, c), rnd(gen, B, c)
    out = torch.zeros(B, 1, n, c, device="cuda", dtype=torch.bfloat16)
    plan = ops.ConvPlan(x, c, ops.pack_linear(w), out, B=B, H=1, W=n, bias=b, bias_img=bi, act=L.ACT_GELU,
                        res=r1, res2=r2)
    plan.run()
    torch.cuda.synchronize()
    ref = F.gelu(F.linear(x.float(), bf(w).float(), b) + bi[:, None, None, :]) + r1.float() + r2.float()
    assert_close_bf16(out, ref, "bias_img + gelu + 2 residuals")


# ---------------------------------------------------------------------------------------------- LinearAttention
@pytest.mark.parametrize("B,H,W,C", [(2, 32, 32, 64), (1, 64, 48, 128), (2, 16, 16, 256)])
def test_linear_attention_block(ops, gen, B, H, W, C):
    """to_qkv (+q softmax) -> context reduce -> fold into to_out -> GEMM + LN + residual, vs the oracle."""
    from daclip_b200 import lib as L
    from oracle import unet_oracle as O
    hw = H * W
    x = rnd(gen, B, C, H, W)
    sd = {"to_qkv.weight": rnd(gen, 384, C, 1, 1, scale=C ** -0.5),
          "to_out.0.weight": rnd(gen, C, 128, 1, 1, scale=128 ** -0.5 * 8),
          "to_out.0.bias": rnd(gen, C, scale=0.1), "to_out.1.g": (1 + 0.1 * rnd(gen, 1, C, 1, 1))}
    xh = nhwc(x)
    sdr = dict(sd)
    sdr["to_qkv.weight"] = bf(sd["to_qkv.weight"]).float()
    ref = O.linear_attention(sdr, "", nchw(xh)) + nchw(xh)

    q = torch.zeros(B, H, W, 128, device="cuda", dtype=torch.bfloat16)
    kv = torch.full((B, 256, H, W), float("nan"), device="cuda", dtype=torch.bfloat16)
    p1 = ops.ConvPlan(xh, C, ops.pack_linear(sd["to_qkv.weight"]), q, B=B, H=H, W=W, epi=L.EPI_QKV, block_n=128,
                      out_planar=kv)
    nchunks = 8
    partial = torch.zeros(B, 4, nchunks, 32 * 34, device="cuda")
    c_pad = ops.choose_block_n(C)[1]
    weff = torch.zeros(B, c_pad, 128, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(B, H, W, C, device="cuda", dtype=torch.bfloat16)
    pw_out = ops.pack_linear(sd["to_out.0.weight"])
    p2 = ops.ConvPlan(q, 128, pw_out, out, B=B, H=H, W=W, epi=L.EPI_LN, bias=sd["to_out.0.bias"],
                      ln_g=sd["to_out.1.g"].reshape(-1).contiguous(), res=xh, per_image_w=True, weight_override=weff)
    p1.run()
    ops.linattn_context(kv, B, hw, nchunks, partial)
    ops.linattn_fold(partial, B, hw, nchunks, sd["to_out.0.weight"].reshape(C, 128).contiguous(), C, c_pad, weff)
    p2.run()
    torch.cuda.synchronize()
    # q check (softmax over head channels * scale)
    q_ref = F.conv2d(nchw(xh), sdr["to_qkv.weight"])[:, :128].reshape(B, 4, 32, hw).softmax(2) * 32 ** -0.5
    q_got = q.float().reshape(B, hw, 4, 32).permute(0, 2, 3, 1)
    assert_close_bf16(q_got, q_ref, "q softmax epilogue")
    kv_ref = F.conv2d(nchw(xh), sdr["to_qkv.weight"])[:, 128:]
    assert_close_bf16(kv, kv_ref, "planar k|v output")
    assert_close_bf16(nchw(out), ref, f"linear attention {H}x{W} C={C}", rel=2 ** -5, abs_=2e-2)


@pytest.mark.parametrize("B,H,W,C", [(3, 32, 32, 64), (2, 24, 40, 128), (2, 16, 16, 256), (5, 72, 56, 64)])
def test_linear_attention_fused_kv_context(ops, gen, B, H, W, C):
    """KVCTX epilogue: k | v are reduced into the per-(image, head) context inside the GEMM epilogue with the
    data-independent shift c_d = ||W_k[d]|| sqrt(C); checked through the fold against softmax_pixels(k) v^T / hw in
    fp32 (module_util.py:170-177), on ragged sizes (masked tile rows) and several images per CTA (flushes)."""
    from daclip_b200 import lib as L
    hw = H * W
    x = rnd(gen, B, C, H, W) * 1.7 + 0.3
    xn = (x - x.mean(1, keepdim=True)) * torch.rsqrt(x.var(1, unbiased=False, keepdim=True) + 1e-5)
    xh = nhwc(xn)
    wkv = rnd(gen, 256, C, scale=C ** -0.5)
    wkv[:128] *= 1.5                                           # a wider softmax than unit-variance logits
    w_out = rnd(gen, C, 128, scale=128 ** -0.5)
    wk = bf(wkv[:128]).float()
    shift = (1.02 * wk.norm(dim=1) * math.sqrt(C))
    assert shift.max().item() <= 40
    ns = ops.ctx_slots(B, H, W, False)
    ctx = torch.full((B, 4, ns, 32 * 34), float("nan"), device="cuda")
    plan = ops.ConvPlan(xh, C, ops.pack_linear(wkv), None, B=B, H=H, W=W, epi=L.EPI_KVCTX, block_n=256,
                        kv_shift=(shift * 1.4426950408889634).contiguous(), ctx_acc=ctx)
    c_pad = ops.choose_block_n(C)[1]
    weff = torch.zeros(B, c_pad, 128, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):                                         # relaunch: the partial records are re-zeroed every time
        plan.run()
    ops.linattn_fold(ctx, B, hw, ns, w_out, C, c_pad, weff)
    torch.cuda.synchronize()
    first = (ctx.clone(), weff.clone())
    plan.run()                                                 # stores in fixed slots, fixed-order merge: bit-reproducible
    ops.linattn_fold(ctx, B, hw, ns, w_out, C, c_pad, weff)
    torch.cuda.synchronize()
    assert torch.equal(ctx, first[0]) and torch.equal(weff, first[1])
    with pytest.raises(L.DacError):                            # too few slots is an error, not a silent overflow
        ops.ConvPlan(xh, C, ops.pack_linear(wkv), None, B=B, H=H, W=W, epi=L.EPI_KVCTX, block_n=256,
                     kv_shift=(shift * 1.4426950408889634).contiguous(), ctx_acc=ctx[:, :, :1].contiguous())
    ctx = ctx.sum(2, keepdim=True)
    kvr = F.conv2d(nchw(xh), bf(wkv).float()[:, :, None, None]).reshape(B, 2, 4, 32, hw)
    k, v = kvr[:, 0].softmax(-1), kvr[:, 1] / hw
    ctx_ref = torch.einsum("bhdn,bhen->bhde", k, v)            # [B, 4, 32, 32]
    got = ctx[:, :, 0, :1024].reshape(B, 4, 32, 32) / (ctx[:, :, 0, 1056:1088, None] * hw)
    scale = ctx_ref.abs().max().item()
    assert (got - ctx_ref).abs().max().item() <= 6e-3 * scale, (got - ctx_ref).abs().max().item() / scale
    weff_ref = torch.einsum("che,bhde->bchd", w_out.reshape(C, 4, 32), ctx_ref).reshape(B, C, 128)
    assert_close_bf16(weff[:, :C], weff_ref, "folded context weight", rel=2 ** -6, abs_=1e-2)


@pytest.mark.parametrize("B,H,W,C", [(3, 32, 32, 64), (2, 16, 40, 128), (5, 64, 48, 64), (1, 8, 16, 128), (40, 16, 8, 64)])
def test_linear_attention_kv_tensor_core_context(ops, gen, B, H, W, C):
    """dac_linattn_kv: k|v GEMM -> exp -> P^T V / P^T 1 as a second tcgen05 GEMM with MN-major operands, context held
    in tensor memory across tiles and flushed per image; checked against softmax_pixels(k) v^T / hw in fp32."""
    hw = H * W
    x = rnd(gen, B, C, H, W) * 1.7 + 0.3
    xn = (x - x.mean(1, keepdim=True)) * torch.rsqrt(x.var(1, unbiased=False, keepdim=True) + 1e-5)
    xh = nhwc(xn)
    wkv = rnd(gen, 256, C, scale=C ** -0.5)
    wkv[:128] *= 1.5
    shift = 1.02 * bf(wkv[:128]).float().norm(dim=1) * math.sqrt(C)
    ctx = torch.full((B, 4, ops.ctx_slots(B, H, W, True), 32 * 34), float("nan"), device="cuda")
    plan = ops.KvPlan(xh, ops.pack_kv_grouped(wkv), (shift * 1.4426950408889634).contiguous(), ctx, B, hw, C)
    for _ in range(2):
        plan.run()
    torch.cuda.synchronize()
    first = ctx.clone()
    plan.run()
    torch.cuda.synchronize()
    assert torch.equal(ctx, first)                             # one record per (CTA, image), plain stores
    ctx = ctx.sum(2, keepdim=True)
    kvr = F.conv2d(nchw(xh), bf(wkv).float()[:, :, None, None]).reshape(B, 2, 4, 32, hw)
    k, v = kvr[:, 0].softmax(-1), kvr[:, 1] / hw
    ctx_ref = torch.einsum("bhdn,bhen->bhde", k, v)
    got = ctx[:, :, 0, :1024].reshape(B, 4, 32, 32) / (ctx[:, :, 0, 1056:1088, None] * hw)
    scale = ctx_ref.abs().max().item()
    assert torch.isfinite(got).all()
    assert (got - ctx_ref).abs().max().item() <= 6e-3 * scale, (got - ctx_ref).abs().max().item() / scale


@pytest.mark.parametrize("B,H,W", [(3, 32, 32), (5, 64, 48), (40, 16, 8), (1, 128, 128), (16, 8, 16)])
def test_linear_attention_kv_in_kernel_prenorm(ops, gen, B, H, W):
    """dac_linattn_kv with prenorm: the RAW 64-channel tensor goes in, every tile is normalised in shared memory (gain-free
    channel LayerNorm, module_util.py:77-97) before the k GEMM, and the kernel accumulates G = P^T xn and S = P^T 1 per
    image (the values are never formed); checked as G W_v^T / (S hw) against LayerNorm -> bf16 -> to_kv ->
    softmax_pixels(k) v^T / hw in fp32 (module_util.py:170-177), and bit-reproducible."""
    C, hw = 64, H * W
    x = rnd(gen, B, C, H, W) * 1.7 + 0.3
    xr = nhwc(x)                                               # what the producing layer stored (bf16)
    xf = nchw(xr)
    xn = bf((xf - xf.mean(1, keepdim=True)) * torch.rsqrt(xf.var(1, unbiased=False, keepdim=True) + 1e-5)).float()
    wkv = rnd(gen, 256, C, scale=C ** -0.5)
    wkv[:128] *= 1.5
    shift = 1.02 * bf(wkv[:128]).float().norm(dim=1) * math.sqrt(C)
    ctx = torch.full((B, 4, ops.ctx_slots(B, H, W, True), ops.KV_G_REC), float("nan"), device="cuda")
    plan = ops.KvPlan(xr, ops.centre_rows(wkv[:128]).to(torch.bfloat16).contiguous(),
                      (shift * 1.4426950408889634).contiguous(), ctx, B, hw, C, prenorm_eps=1e-5)
    for _ in range(2):
        plan.run()
    torch.cuda.synchronize()
    first = ctx.clone()
    plan.run()
    torch.cuda.synchronize()
    assert torch.equal(ctx, first)
    assert torch.equal(nhwc(x), xr)                            # the input tensor is not modified (normalised in smem only)
    written = ~torch.isnan(ctx).any(-1)                        # records no CTA owns stay untouched (and are never read)
    assert written[:, :, 0].all() and (written == written[:, :1]).all()
    rec = torch.where(written[..., None], ctx, torch.zeros_like(ctx)).sum(2)     # [B, 4, 2080]
    G, S = rec[..., :2048].reshape(B, 4, 32, 64), rec[..., 2048:2080]
    wv = bf(ops.centre_rows(wkv[128:])).float().reshape(4, 32, C)     # centred: the per-pixel mean drops out of G W_v^T
    wv = wv - wv.mean(2, keepdim=True)
    got = torch.einsum("bhdc,hec->bhde", G, wv) / (S[..., None] * hw)
    kvr = F.conv2d(xn, bf(wkv).float()[:, :, None, None]).reshape(B, 2, 4, 32, hw)
    k, v = kvr[:, 0].softmax(-1), kvr[:, 1] / hw
    ctx_ref = torch.einsum("bhdn,bhen->bhde", k, v)
    scale = ctx_ref.abs().max().item()
    assert torch.isfinite(got).all()
    assert (got - ctx_ref).abs().max().item() <= 6e-3 * scale, (got - ctx_ref).abs().max().item() / scale


@pytest.mark.parametrize("B,H,W,C", [(3, 32, 32, 64), (2, 16, 40, 128), (5, 64, 48, 64), (1, 8, 16, 128)])
def test_linear_attention_chained_q_out(ops, gen, B, H, W, C):
    """Whole LinearAttention block on the fused path: KVCTX context -> fold -> chained kernel (to_q, head softmax,
    W_eff q, LayerNorm, + residual), vs the oracle block (module_util.py:157-185) in fp32."""
    from daclip_b200 import lib as L
    from oracle import unet_oracle as O
    hw = H * W
    x = rnd(gen, B, C, H, W) * 1.3 + 0.2
    xn = (x - x.mean(1, keepdim=True)) * torch.rsqrt(x.var(1, unbiased=False, keepdim=True) + 1e-5)
    xh, rh = nhwc(xn), nhwc(rnd(gen, B, C, H, W))
    sd = {"to_qkv.weight": rnd(gen, 384, C, 1, 1, scale=C ** -0.5),
          "to_out.0.weight": rnd(gen, C, 128, 1, 1, scale=128 ** -0.5 * 8),
          "to_out.0.bias": rnd(gen, C, scale=0.1), "to_out.1.g": (1 + 0.1 * rnd(gen, 1, C, 1, 1))}
    sdr = dict(sd)
    sdr["to_qkv.weight"] = bf(sd["to_qkv.weight"]).float()
    ref = O.linear_attention(sdr, "", nchw(xh)) + nchw(rh)
    wqkv = sd["to_qkv.weight"].reshape(384, C)
    shift = 1.02 * bf(wqkv[128:256]).float().norm(dim=1) * math.sqrt(C)
    ns = ops.ctx_slots(B, H, W, False)
    ctx = torch.zeros(B, 4, ns, 32 * 34, device="cuda")
    c_pad = ops.choose_block_n(C)[1]
    weff = torch.zeros(B, c_pad, 128, device="cuda", dtype=torch.bfloat16)
    out = torch.full((B, H, W, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    pkv = ops.ConvPlan(xh, C, ops.pack_linear(wqkv[128:].contiguous()), None, B=B, H=H, W=W, epi=L.EPI_KVCTX,
                       block_n=256, kv_shift=(shift * 1.4426950408889634).contiguous(), ctx_acc=ctx)
    wq = ops.pack_linear(wqkv[:128].contiguous())
    pq = ops.QoutPlan(xh, wq.w, weff, rh, out, sd["to_out.0.bias"], sd["to_out.1.g"].reshape(-1).contiguous(), 1e-5,
                      B, hw, C)
    for _ in range(2):
        pkv.run()
        ops.linattn_fold(ctx, B, hw, ns, sd["to_out.0.weight"].reshape(C, 128).contiguous(), C, c_pad, weff)
        pq.run()
    torch.cuda.synchronize()
    assert_close_bf16(nchw(out), ref, f"chained linear attention {H}x{W} C={C}", rel=2 ** -5, abs_=2e-2)


@pytest.mark.parametrize("B,H,W,C,kv_tc", [(3, 32, 32, 64, True), (2, 16, 40, 128, False), (5, 64, 48, 64, True),
                                            (1, 8, 16, 128, False), (2, 16, 24, 64, False)])
def test_linear_attention_folded_prenorm(ops, gen, B, H, W, C, kv_tc):
    """PreNorm folded into the fused LinearAttention kernels: the k|v kernel (tcgen05 context or KVCTX epilogue) and the
    chained q / to_out kernel take the RAW tensor plus the per-pixel {mean, rstd} its producer wrote and gain-folded
    weights with their column sums; checked against PreNorm(LinearAttention) of the oracle on the same tensor."""
    from daclip_b200 import lib as L
    from oracle import unet_oracle as O
    hw = H * W
    x = rnd(gen, B, C, H, W) * 1.7 + 0.4
    xh = nhwc(x)                                                 # bf16 NHWC: what the producing ResBlock stored
    xf = nchw(xh)
    g = 1 + 0.2 * rnd(gen, C)
    sd = {"to_qkv.weight": rnd(gen, 384, C, 1, 1, scale=C ** -0.5),
          "to_out.0.weight": rnd(gen, C, 128, 1, 1, scale=128 ** -0.5 * 8),
          "to_out.0.bias": rnd(gen, C, scale=0.1), "to_out.1.g": (1 + 0.1 * rnd(gen, 1, C, 1, 1))}
    wf = sd["to_qkv.weight"].reshape(384, C) * g[None, :]        # W' = W diag(g)
    sdr = dict(sd)
    sdr["to_qkv.weight"] = bf(wf).float().reshape(384, C, 1, 1)
    xn = (xf - xf.mean(1, keepdim=True)) * torch.rsqrt(xf.var(1, unbiased=False, keepdim=True) + 1e-5)
    ref = O.linear_attention(sdr, "", xn) + xf
    # the producer's statistics (moments of the stored bf16 row, E[x^2] - mean^2)
    rows = xh.float().reshape(B * hw, C)
    mean = rows.mean(1)
    stats = torch.stack([mean, torch.rsqrt((rows * rows).mean(1) - mean * mean + 1e-5)], 1).contiguous()
    colsum = bf(wf).float().sum(1)
    shift = 1.02 * bf(wf[128:256]).float().norm(dim=1) * math.sqrt(C)
    ns = ops.ctx_slots(B, H, W, kv_tc)
    ctx = torch.zeros(B, 4, ns, 32 * 34, device="cuda")
    c_pad = ops.choose_block_n(C)[1]
    weff = torch.zeros(B, c_pad, 128, device="cuda", dtype=torch.bfloat16)
    out = torch.full((B, H, W, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    sh = (shift * 1.4426950408889634).contiguous()
    if kv_tc:
        grouped = ops.pack_kv_grouped(wf[128:])
        pkv = ops.KvPlan(xh, grouped, sh, ctx, B, hw, C, ln_stats=stats, ln_colsum=grouped.float().sum(1).contiguous())
    else:
        pkv = ops.ConvPlan(xh, C, ops.pack_linear(wf[128:].contiguous()), None, B=B, H=H, W=W, epi=L.EPI_KVCTX,
                           block_n=256, kv_shift=sh, ctx_acc=ctx, ln_stats=stats, ln_colsum=colsum[128:].contiguous())
    pq = ops.QoutPlan(xh, ops.pack_linear(wf[:128].contiguous()).w, weff, xh, out, sd["to_out.0.bias"],
                      sd["to_out.1.g"].reshape(-1).contiguous(), 1e-5, B, hw, C, ln_stats=stats,
                      ln_colsum=colsum[:128].contiguous())
    for _ in range(2):
        pkv.run()
        ops.linattn_fold(ctx, B, hw, ns, sd["to_out.0.weight"].reshape(C, 128).contiguous(), C, c_pad, weff)
        pq.run()
    torch.cuda.synchronize()
    assert_close_bf16(nchw(out), ref, f"folded PreNorm linear attention {H}x{W} C={C}", rel=2 ** -5, abs_=2e-2)


@pytest.mark.parametrize("bound", [True, False])
@pytest.mark.parametrize("B,H,W", [(3, 32, 32), (5, 64, 48), (1, 8, 16), (40, 16, 8), (2, 128, 128)])
def test_linear_attention_in_kernel_prenorm(ops, gen, B, H, W, bound):
    """PreNorm inside the fused 64-channel LinearAttention kernels (dac_linattn_kv / dac_linattn_qout with prenorm): both
    take the RAW tensor, normalise each tile in shared memory, and the q-out kernel reuses the raw tile as the residual;
    checked against PreNorm(LinearAttention) + x of the oracle (module_util.py:89-97,157-185) and for bit-reproducibility."""
    from oracle import unet_oracle as O
    C, hw = 64, H * W
    x = rnd(gen, B, C, H, W) * 1.7 + 0.4
    xh = nhwc(x)                                                 # bf16 NHWC: what the producing ResBlock stored
    xf = nchw(xh)
    g = 1 + 0.2 * rnd(gen, C)
    sd = {"to_qkv.weight": rnd(gen, 384, C, 1, 1, scale=C ** -0.5),
          "to_out.0.weight": rnd(gen, C, 128, 1, 1, scale=128 ** -0.5 * 8),
          "to_out.0.bias": rnd(gen, C, scale=0.1), "to_out.1.g": (1 + 0.1 * rnd(gen, 1, C, 1, 1))}
    wf = sd["to_qkv.weight"].reshape(384, C) * g[None, :]        # W' = W diag(g)
    sdr = dict(sd)
    sdr["to_qkv.weight"] = bf(wf).float().reshape(384, C, 1, 1)
    xn = (xf - xf.mean(1, keepdim=True)) * torch.rsqrt(xf.var(1, unbiased=False, keepdim=True) + 1e-5)
    ref = O.linear_attention(sdr, "", xn) + xf
    shift = 1.02 * bf(wf[128:256]).float().norm(dim=1) * math.sqrt(C)
    ns = ops.ctx_slots(B, H, W, True)
    ctx = torch.zeros(B, 4, ns, ops.KV_G_REC, device="cuda")
    c_pad = ops.choose_block_n(C)[1]
    weff = torch.zeros(B, c_pad, 128, device="cuda", dtype=torch.bfloat16)
    out = torch.full((B, H, W, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    pkv = ops.KvPlan(xh, ops.centre_rows(wf[128:256]).to(torch.bfloat16).contiguous(),
                     (shift * 1.4426950408889634).contiguous(), ctx, B, hw, C, prenorm_eps=1e-5)
    m_fold = ops.kv_fold_matrix(sd["to_out.0.weight"].reshape(C, 128), wf[256:])
    wqc = ops.centre_rows(wf[:128]).contiguous()
    qsh = (1.02 * bf(wqc).float().norm(dim=1) * math.sqrt(C) * 1.4426950408889634).contiguous() if bound else None
    pq = ops.QoutPlan(xh, ops.pack_linear(wqc).w, weff, xh, out, sd["to_out.0.bias"],
                      sd["to_out.1.g"].reshape(-1).contiguous(), 1e-5, B, hw, C, prenorm_eps=1e-5, q_shift=qsh)
    outs = []
    for _ in range(3):
        out.fill_(float("nan"))
        pkv.run()
        ops.linattn_fold_g(ctx, B, hw, ns, m_fold, C, c_pad, weff)
        pq.run()
        torch.cuda.synchronize()
        outs.append(out.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
    assert torch.equal(nhwc(x), xh)                              # the input is not modified
    assert_close_bf16(nchw(out), ref, f"in-kernel PreNorm linear attention {H}x{W}", rel=2 ** -5, abs_=2e-2)


@pytest.mark.parametrize("B,H,W,C", [(2, 32, 32, 64), (1, 16, 48, 128), (1, 16, 16, 256)])
def test_prenorm_folded_into_qkv(ops, gen, B, H, W, C):
    """PreNorm (channel LayerNorm, gain only) folded around to_qkv: the producer writes per-pixel {mean, rstd} of its
    bf16 output, the QKV conv runs on the raw tensor with W diag(g) and finishes the normalisation in its epilogue."""
    from daclip_b200 import lib as L
    x = rnd(gen, B, C, H, W) * 2 + 0.7
    g = 1 + 0.2 * rnd(gen, C)
    wqkv = rnd(gen, 384, C, scale=C ** -0.5)
    xh = nhwc(x)
    # producer: an identity 1x1 conv standing in for the ResBlock tail; writes y = x (bf16) and the row statistics
    y = torch.zeros(B, H, W, C, device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(B * H * W, 2, device="cuda")
    ops.ConvPlan(xh, C, ops.pack_linear(torch.eye(C, device="cuda")), y, B=B, H=H, W=W, stats_out=stats).run()
    wf = wqkv * g[None, :]
    q = torch.zeros(B, H, W, 128, device="cuda", dtype=torch.bfloat16)
    kv = torch.zeros(B, 256, H, W, device="cuda", dtype=torch.bfloat16)
    ops.ConvPlan(y, C, ops.pack_linear(wf), q, B=B, H=H, W=W, epi=L.EPI_QKV, block_n=128, out_planar=kv,
                 ln_stats=stats, ln_colsum=bf(wf).float().sum(1).contiguous()).run()
    torch.cuda.synchronize()
    assert torch.equal(y, xh)
    xf = nchw(xh)
    mean, var = xf.mean(1), xf.var(1, unbiased=False)
    assert (stats[:, 0].reshape(B, H, W) - mean).abs().max().item() < 1e-4
    assert (stats[:, 1].reshape(B, H, W) * torch.sqrt(var + 1e-5) - 1).abs().max().item() < 1e-3
    xn = (xf - mean[:, None]) * torch.rsqrt(var[:, None] + 1e-5) * g[None, :, None, None]
    ref = F.conv2d(xn, wqkv[:, :, None, None])
    assert_close_bf16(kv, ref[:, 128:], "folded PreNorm k|v", rel=2 ** -6, abs_=4e-3)
    q_ref = ref[:, :128].reshape(B, 4, 32, H * W).softmax(2) * 32 ** -0.5
    assert_close_bf16(q.float().reshape(B, H * W, 4, 32).permute(0, 2, 3, 1), q_ref, "folded PreNorm q", rel=2 ** -5,
                      abs_=4e-3)


# ---------------------------------------------------------------------------------------------- CTA-pair mode
@pytest.mark.parametrize("kind,B,H,W,cin,cout", [
    ("1x1", 2, 32, 32, 512, 512), ("1x1", 1, 1, 1280, 768, 2304), ("f32", 1, 1, 1280, 768, 768),
    ("3x3", 2, 32, 32, 256, 256), ("3x3", 4, 64, 64, 128, 128), ("geglu", 2, 32, 32, 512, 4096),
    ("pair", 2, 64, 64, 64, 64), ("pair", 2, 64, 64, 128, 64), ("pair_res", 2, 64, 64, 64, 64),
    ("pair_skip", 2, 64, 64, 64, 64), ("skip", 2, 64, 64, 128, 128), ("down", 2, 64, 64, 64, 128)])
def test_cta_pair_mode_is_bit_identical(ops, gen, monkeypatch, kind, B, H, W, cin, cout):
    """conv_igemm_kernel<..., CTA2 = true> (tcgen05.mma.cta_group::2 on 2-CTA clusters: every CTA loads its own activation tile
    and half of the weight rows, the leader issues M = 256 MMAs for both) against the 1-CTA build of the same layer on the same
    inputs: the accumulation order of every output element is the same, so the outputs must agree bit for bit - streamed
    weights (1x1 / Linear / 3x3 / GEGLU / fp32 residual stream), the pixel-pair layers with their per-rank weight layouts,
    fused skip convs; and the 1-CTA result against the fp32 reference of the op (F.conv2d / F.linear)."""
    x = bf(rnd(gen, B, H, W, cin))
    outs = []
    for cta2 in ("0", "1"):
        monkeypatch.setenv("DAC_CTA2", cta2)
        ref = None
        if kind in ("1x1", "f32"):
            w, b = rnd(gen.manual_seed(11), cout, cin) * cin ** -0.5, rnd(gen, cout)
            if kind == "1x1":
                out = torch.zeros(B, H, W, cout, device="cuda", dtype=torch.bfloat16)
                plan = ops.ConvPlan(x, cin, ops.pack_linear(w), out, B=B, H=H, W=W, bias=b, act=ops.L.ACT_GELU)
                ref = F.gelu(F.linear(x.float(), bf(w).float(), b))
            else:
                out = rnd(gen, B, H, W, cout)
                ref = out + F.linear(x.float(), bf(w).float(), b)
                plan = ops.ConvPlan(x, cin, ops.pack_linear(w), None, B=B, H=H, W=W, bias=b, res_f32=out, out_f32=out)
        elif kind == "geglu":
            w, b = rnd(gen.manual_seed(11), cout, cin) * cin ** -0.5, rnd(gen, cout)
            pw, bp = ops.pack_geglu(w, b)
            out = torch.zeros(B, H, W, cout // 2, device="cuda", dtype=torch.bfloat16)
            plan = ops.ConvPlan(x, cin, pw, out, B=B, H=H, W=W, epi=ops.L.EPI_GEGLU, bias=bp, block_n=256)
            y = F.linear(x.float(), bf(w).float(), b)
            ref = y[..., :cout // 2] * F.gelu(y[..., cout // 2:])
        elif kind == "down":
            w, b = rnd(gen.manual_seed(11), cout, cin, 4, 4) * (16 * cin) ** -0.5, rnd(gen, cout)
            out = torch.zeros(B, H // 2, W // 2, cout, device="cuda", dtype=torch.bfloat16)
            plan = ops.ConvPlan(x, cin, ops.pack_conv(w, stride=2, pad=1), out, B=B, H=H, W=W, bias=b)
            ref = F.conv2d(x.float().permute(0, 3, 1, 2), bf(w).float(), b, stride=2, padding=1).permute(0, 2, 3, 1)
        else:
            w = rnd(gen.manual_seed(11), cout, cin, 3, 3) * (9 * cin) ** -0.5
            film = rnd(gen, B, 2 * cout) * 0.1
            res = bf(rnd(gen, B, H, W, cout))
            out = torch.zeros(B, H, W, cout, device="cuda", dtype=torch.bfloat16)
            conv = F.conv2d(x.float().permute(0, 3, 1, 2), bf(w).float(), padding=1).permute(0, 2, 3, 1)
            if kind == "3x3":
                plan = ops.ConvPlan(x, cin, ops.pack_conv(w), out, B=B, H=H, W=W, act=ops.L.ACT_SILU, film=film)
                ref = F.silu(conv * (film[:, None, None, :cout] + 1) + film[:, None, None, cout:])
            elif kind == "pair":
                a, s1 = x[..., :64].contiguous(), (x[..., 64:].contiguous() if cin == 128 else None)
                plan = ops.PairConvPlan(a, ops.pack_conv_pair(w), out, B=B, H=H, W=W, src1=s1, act=ops.L.ACT_SILU, film=film)
                ref = F.silu(conv * (film[:, None, None, :cout] + 1) + film[:, None, None, cout:])
            elif kind == "pair_res":
                plan = ops.PairConvPlan(x, ops.pack_conv_pair(w), out, B=B, H=H, W=W, act=ops.L.ACT_SILU, res=res)
                ref = F.silu(conv) + res.float()
            else:
                rc = 128 if kind == "pair_skip" else cout + cout // 2
                wr = rnd(gen, cout, rc) * rc ** -0.5
                r0, r1 = bf(rnd(gen, B, H, W, cout)), bf(rnd(gen, B, H, W, rc - cout))
                skip = F.linear(torch.cat([r0, r1], -1).float(), bf(wr).float())
                ref = F.silu(conv) + skip
                if kind == "pair_skip":
                    plan = ops.PairConvPlan(x, ops.pack_conv_pair(w), out, B=B, H=H, W=W, act=ops.L.ACT_SILU, rsrc0=r0,
                                            rsrc1=r1, rweight=ops.pack_linear(wr))
                else:
                    plan = ops.ConvPlan(x, cin, ops.pack_conv(w), out, B=B, H=H, W=W, act=ops.L.ACT_SILU, rsrc0=r0, rc0=cout,
                                        rsrc1=r1, rc1=rc - cout, rweight=ops.pack_linear(wr))
        plan.run()
        torch.cuda.synchronize()
        outs.append((out.clone(), ref))
    assert torch.equal(outs[0][0], outs[1][0]), (outs[0][0].float() - outs[1][0].float()).abs().max().item()
    if outs[0][0].dtype == torch.bfloat16:
        assert_close_bf16(outs[0][0], outs[0][1], f"cta-pair case {kind}", rel=2 ** -6, abs_=6e-3)
    else:
        assert (outs[0][0] - outs[0][1]).abs().max().item() <= 2e-2 * outs[0][1].abs().max().item()


# ---------------------------------------------------------------------------------------------- attention
@pytest.mark.parametrize("B,n,heads", [(2, 1024, 8), (1, 4096, 16), (1, 200, 4), (1, 128, 2), (3, 384, 6), (5, 1024, 16)])
def test_flash_attention_d32(ops, gen, B, n, heads):
    d = 32
    qkv = bf(rnd(gen, B, n, 3 * heads * d))
    out = torch.zeros(B, n, heads * d, device="cuda", dtype=torch.bfloat16)
    ops.attention(qkv, out, B, n, heads, d)
    torch.cuda.synchronize()
    q, k, v = [t.reshape(B, n, heads, d).transpose(1, 2) for t in qkv.float().chunk(3, dim=-1)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, n, heads * d)
    assert_close_bf16(out, ref, f"flash d32 n={n}", rel=2 ** -6, abs_=4e-3)


@pytest.mark.parametrize("B,n,heads", [(1, 1024, 2), (2, 2048, 4)])
def test_flash_attention_d32_growing_scores(ops, gen, B, n, heads):
    """The tcgen05 attention kernel keeps O in tensor memory and rescales a row only when its running max grows by more
    than 2^8 (attention_tc2.cu).  Random scores never trigger that path: here the key norms step up x4 every 256 keys, so
    the row max jumps by far more than the threshold from block to block (and some rows shrink instead, which must NOT
    rescale), against the fp32 softmax(QK^T / sqrt(d)) V of attention.py:178-192."""
    d = 32
    qkv = rnd(gen, B, n, 3 * heads * d)
    step = 4.0 ** torch.div(torch.arange(n, device="cuda"), 256, rounding_mode="floor").float()
    kcols = slice(heads * d, 2 * heads * d)
    qkv[:, :, kcols] = qkv[:, :, kcols] * step[None, :, None] * 0.5
    qkv = bf(qkv)
    out = torch.zeros(B, n, heads * d, device="cuda", dtype=torch.bfloat16)
    ops.attention(qkv, out, B, n, heads, d)
    torch.cuda.synchronize()
    q, k, v = [t.reshape(B, n, heads, d).transpose(1, 2) for t in qkv.float().chunk(3, dim=-1)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, n, heads * d)
    assert torch.isfinite(out.float()).all()
    assert_close_bf16(out, ref, f"flash d32 growing scores n={n}", rel=2 ** -6, abs_=6e-3)


def test_flash_attention_d32_polynomial_exp(ops, gen, monkeypatch):
    """A quarter of the exponentials of attn_tc2_kernel run as a polynomial on the FMA pipe (ex2_poly2, attention_tc2.cu).
    Scores far below the row max (here down to -600 in log2 units: the argument clamp) must give weight zero, not a wrapped
    exponent; the all-MUFU build (DAC_ATTN_POLY=0) is the second witness next to the fp32 softmax of attention.py:178-192."""
    B, n, heads, d = 2, 1024, 4, 32
    qkv = rnd(gen, B, n, 3 * heads * d)
    kcols = slice(heads * d, 2 * heads * d)
    qkv[:, ::7, kcols] *= 24.0                      # a few dominant keys: every other score sits far below the max
    qkv = bf(qkv)
    outs = []
    for poly in ("1", "0"):
        monkeypatch.setenv("DAC_ATTN_POLY", poly)
        out = torch.zeros(B, n, heads * d, device="cuda", dtype=torch.bfloat16)
        ops.attention(qkv, out, B, n, heads, d)
        torch.cuda.synchronize()
        outs.append(out)
    q, k, v = [t.reshape(B, n, heads, d).transpose(1, 2) for t in qkv.float().chunk(3, dim=-1)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, n, heads * d)
    for out, tag in zip(outs, ("polynomial", "MUFU")):
        assert torch.isfinite(out.float()).all()
        assert_close_bf16(out, ref, f"flash d32 {tag}", rel=2 ** -6, abs_=6e-3)
    assert (outs[0].float() - outs[1].float()).abs().max().item() <= 2 ** -6 * ref.abs().max().item()


def test_small_attention_d64(ops, gen):
    B, n, heads, d = 3, 50, 12, 64
    qkv = bf(rnd(gen, B, n, 3 * heads * d))
    out = torch.zeros(B, n, heads * d, device="cuda", dtype=torch.bfloat16)
    ops.attention(qkv, out, B, n, heads, d)
    torch.cuda.synchronize()
    q, k, v = [t.reshape(B, n, heads, d).transpose(1, 2) for t in qkv.float().chunk(3, dim=-1)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, n, heads * d)
    assert_close_bf16(out, ref, "small attn d64")


@pytest.mark.parametrize("B,n,heads,causal", [(3, 50, 12, False), (300, 50, 12, False), (2, 257, 16, False), (5, 128, 2, False),
                                              (1, 1, 3, False), (2, 400, 4, False), (10, 77, 8, True), (3, 128, 12, True),
                                              (2, 5, 1, True)])
def test_attention_d64_tcgen05(ops, gen, B, n, heads, causal):
    """attn_vit_kernel (tcgen05): the ViT blocks' attention (50 / 257 tokens, transformer.py:219-230) and, causal, the text
    tower's (77 tokens, model.py:237-249) - unused tile rows, several key blocks, many items per CTA, growing scores."""
    d = 64
    qkv = rnd(gen, B, n, 3 * heads * d)
    qkv[:, :, heads * d:2 * heads * d] *= (1.0 + 3.0 * torch.arange(n, device="cuda").float() / max(n, 2))[None, :, None]
    qkv = bf(qkv)
    out = torch.full((B, n, heads * d), float("nan"), device="cuda", dtype=torch.bfloat16)
    guard = out.clone()
    (ops.attention_causal if causal else ops.attention)(qkv, out, B, n, heads, d)
    torch.cuda.synchronize()
    q, k, v = [t.reshape(B, n, heads, d).transpose(1, 2) for t in qkv.float().chunk(3, dim=-1)]
    ref = F.scaled_dot_product_attention(q, k, v, is_causal=causal).transpose(1, 2).reshape(B, n, heads * d)
    assert torch.isfinite(out.float()).all()
    assert_close_bf16(out, ref, f"attention d64 n={n} causal={causal}", rel=2 ** -6, abs_=6e-3)
    out2 = guard.clone()
    (ops.attention_causal if causal else ops.attention)(qkv, out2, B, n, heads, d)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)                                # bit-reproducible


# ---------------------------------------------------------------------------------------------- norms
@pytest.mark.parametrize("rows,c,affine", [(1000, 64, "g"), (4096, 512, "wb"), (350, 768, "wb")])
def test_layernorm_rows(ops, gen, rows, c, affine):
    x = bf(rnd(gen, rows, c) * 2 + 0.5)
    w = 1 + 0.1 * rnd(gen, c)
    b = rnd(gen, c) if affine == "wb" else None
    out = torch.zeros_like(x)
    ops.layernorm_rows(x, out, rows, c, w, b, 1e-5)
    torch.cuda.synchronize()
    ref = F.layer_norm(x.float(), (c,), w, b, 1e-5)
    assert_close_bf16(out, ref, "layernorm rows")


def test_fp32_residual_stream(ops, gen):
    """ViT block pattern: x(fp32) -> LN -> bf16 -> GEMM + bias + fp32 residual (+ bf16 control) -> fp32 and bf16."""
    rows, c = 350, 768
    x = rnd(gen, 1, 1, rows, c) * 2
    lw, lb = 1 + 0.1 * rnd(gen, c), rnd(gen, c)
    n = torch.zeros(1, 1, rows, c, device="cuda", dtype=torch.bfloat16)
    ops.layernorm_rows_f32(x, n, rows, c, lw, lb, 1e-5)
    torch.cuda.synchronize()
    assert_close_bf16(n, F.layer_norm(x, (c,), lw, lb, 1e-5), "layernorm f32-in")
    w, b = rnd(gen, c, c, scale=c ** -0.5), rnd(gen, c)
    ctrl = bf(rnd(gen, 1, 1, rows, c))
    y32 = torch.zeros(1, 1, rows, c, device="cuda")
    y16 = torch.zeros(1, 1, rows, c, device="cuda", dtype=torch.bfloat16)
    plan = ops.ConvPlan(n, c, ops.pack_linear(w), y16, B=1, H=1, W=rows, bias=b, res_f32=x, res2=ctrl, out_f32=y32)
    plan.run()
    torch.cuda.synchronize()
    ref = F.linear(n.float(), bf(w).float(), b) + x + ctrl.float()
    assert (y32 - ref).abs().max().item() < 2e-3 * ref.abs().max().item()
    assert_close_bf16(y16, ref, "bf16 copy of fp32 stream")


@pytest.mark.parametrize("B,hw,c", [(2, 1024, 256), (3, 4096, 512), (1, 36, 256)])
def test_groupnorm(ops, gen, B, hw, c):
    x = bf(rnd(gen, B, hw, c) + 0.3)
    w, b = 1 + 0.1 * rnd(gen, c), rnd(gen, c)
    out, stats = torch.zeros_like(x), torch.full((B * 16 * 64,), float("nan"), device="cuda")
    ops.groupnorm_nhwc(x, out, B, hw, c, w, b, stats)
    torch.cuda.synchronize()
    ref = F.group_norm(x.float().transpose(1, 2), 32, w, b, 1e-6).transpose(1, 2)
    assert_close_bf16(out, ref, "groupnorm")


@pytest.mark.parametrize("rows,c", [(1000, 512), (333, 64), (77, 1024)])
def test_layernorm_rows_with_residual(ops, gen, rows, c):
    """dac_layernorm_rows_res: LayerNorm(in) * g + res - the to_out LayerNorm and the Residual wrapper of a LinearAttention
    wider than 256 channels (module_util.py:27-33,168,185)."""
    x, res = bf(rnd(gen, rows, c) * 3 + 0.5), bf(rnd(gen, rows, c))
    g = 1 + 0.2 * rnd(gen, c)
    out = torch.zeros_like(x)
    ops.layernorm_rows_res(x, res, out, rows, c, g, None, 1e-5)
    torch.cuda.synchronize()
    xf = x.float()
    ref = (xf - xf.mean(-1, keepdim=True)) * torch.rsqrt(xf.var(-1, unbiased=False, keepdim=True) + 1e-5) * g + res.float()
    assert_close_bf16(out, ref, "layernorm + residual")


@pytest.mark.parametrize("B,hw,c", [(2, 1024, 256), (3, 1024, 512), (1, 100, 512), (2, 64, 768)])
def test_prenorm_groupnorm_fused(ops, gen, B, hw, c):
    """dac_prenorm_groupnorm_nhwc: channel LayerNorm with gain (module_util.py:77-97) and GroupNorm(32, eps 1e-6) of its
    output (attention.py:76-77,251) - the normalised rows and the GroupNorm output against fp32 torch, and the fused
    statistics against the two-kernel path."""
    x = bf(rnd(gen, B, hw, c) * (1 + torch.arange(c, device="cuda") % 5)[None, None, :] + 0.7)
    g = 1 + 0.2 * rnd(gen, c)
    w, b = 1 + 0.1 * rnd(gen, c), rnd(gen, c)
    xn, out = torch.zeros_like(x), torch.zeros_like(x)
    stats = torch.full((B * 16 * 64,), float("nan"), device="cuda")
    ops.prenorm_groupnorm_nhwc(x, xn, out, B, hw, c, g, w, b, stats)
    torch.cuda.synchronize()
    xf = x.float()
    ln = (xf - xf.mean(-1, keepdim=True)) * torch.rsqrt(xf.var(-1, unbiased=False, keepdim=True) + 1e-5) * g
    assert_close_bf16(xn, ln, "fused prenorm")
    ref = F.group_norm(xn.float().transpose(1, 2), 32, w, b, 1e-6).transpose(1, 2)
    assert_close_bf16(out, ref, "fused prenorm -> groupnorm")
    xn2, out2 = torch.zeros_like(x), torch.zeros_like(x)
    ops.layernorm_rows(x, xn2, B * hw, c, g, None, 1e-5)
    ops.groupnorm_nhwc(xn2, out2, B, hw, c, w, b, stats)
    torch.cuda.synchronize()
    assert torch.equal(xn, xn2)
    assert (out.float() - out2.float()).abs().max().item() <= 2 ** -7 * ref.abs().max().item()


# ---------------------------------------------------------------------------------------------- SDE updates
@pytest.mark.parametrize("mode", ["sde", "posterior", "ode"])
def test_sde_step_bit_exact(cuda, mode):
    from daclip_b200.sde import IRSDE
    from oracle import sde_oracle as S
    sched = S.Schedule(50, 100, "cosine", 0.005)
    sde = IRSDE(50, T=100, schedule="cosine", eps=0.005, device=cuda)
    g = torch.Generator().manual_seed(7)
    shape = (2, 3, 37, 53)  # odd sizes: exercises the scalar tail
    x, mu, n, e = [torch.randn(shape, generator=g) for _ in range(4)]
    sde.set_mu(mu.cuda())
    for t in (100, 57, 2, 1):
        if mode == "sde":
            ref = S.sde_step(sched, x, mu, n, e, t)
            got = sde.reverse_sde_step(x.cuda(), sde.get_score_from_noise(n.cuda(), t), t, eps=e.cuda())
        elif mode == "posterior":
            ref = S.posterior_step(sched, x, mu, n, e, t)
            got = sde.reverse_posterior_step(x.cuda(), n.cuda(), t, eps=e.cuda())
        else:
            ref = S.ode_step(sched, x, mu, n, t)
            got = sde.reverse_ode_step(x.cuda(), sde.get_score_from_noise(n.cuda(), t), t)
        diff = (got.cpu() - ref).abs().max().item()
        assert diff <= 2e-7 * max(1.0, ref.abs().max().item()), f"{mode} t={t}: max diff {diff}"


def test_device_driven_step_tick_philox_and_injected_noise(ops, cuda):
    """dac_loop_tick + dac_sde_step_dev: the step counter, time and coefficients come from device tables; the noise is
    either row `step` of a pre-generated tensor (bit-identical to dac_sde_step with the same operands) or Philox normals
    generated in the kernel (mean 0, variance 1, different per step, reproducible per seed)."""
    from daclip_b200.sde import IRSDE
    sde = IRSDE(50, T=100, schedule="cosine", eps=0.005, device=cuda)
    g = torch.Generator(device="cuda").manual_seed(11)
    shape, T = (2, 3, 64, 64), 5
    x, mu, n = [torch.randn(shape, device="cuda", generator=g) for _ in range(3)]
    eps = torch.randn((T,) + shape, device="cuda", generator=g)
    ts = [100, 57, 33, 2, 1]
    coefs = [list(sde._posterior_coef(t)) + [0.0] * 3 for t in ts]
    t_tab = torch.tensor([float(t) for t in ts], device="cuda")
    c_tab = torch.tensor(coefs, device="cuda", dtype=torch.float32)
    t_dev, c_dev = torch.zeros(1, device="cuda"), torch.zeros(8, device="cuda")
    state = torch.tensor([0, 0, eps.data_ptr(), 0], device="cuda", dtype=torch.int64)
    for i, t in enumerate(ts):
        ops.loop_tick(state, t_tab, c_tab, t_dev, c_dev)
        out = torch.empty_like(x)
        ops.sde_step_dev(1, x, mu, n, out, c_dev, state)
        ref = torch.empty_like(x)
        ops.sde_step(1, x, mu, n, eps[i].contiguous(), ref, sde._posterior_coef(t))
        torch.cuda.synchronize()
        assert state[:2].tolist() == [i + 1, i] and t_dev.item() == float(t)
        assert torch.equal(out, ref), f"step {i}"
    # Philox: x = mu = net = 0 and {term1, term2, std, exp, sigma_bar} = {0, 0, 1, 1, 0}: out is the noise itself
    z = torch.zeros(4, 3, 256, 256, device="cuda")
    c_tab = torch.tensor([[0, 0, 1, 1, 0, 0, 0, 0]] * 3, device="cuda", dtype=torch.float32)
    draws = {}
    for seed in (1234, 1234, 99):
        state = torch.tensor([0, 0, 0, seed], device="cuda", dtype=torch.int64)
        outs = []
        for i in range(3):
            ops.loop_tick(state, t_tab, c_tab, t_dev, c_dev)
            o = torch.empty_like(z)
            ops.sde_step_dev(1, z, z, z, o, c_dev, state)
            outs.append(o)
        torch.cuda.synchronize()
        if seed in draws:
            assert all(torch.equal(a, b) for a, b in zip(draws[seed], outs))      # reproducible per seed
        draws[seed] = outs
    a = draws[1234]
    for o in a:
        assert abs(o.mean().item()) < 5e-3 and abs(o.var().item() - 1.0) < 1e-2
        assert abs((o ** 4).mean().item() - 3.0) < 0.1                            # Gaussian kurtosis
    assert not torch.equal(a[0], a[1]) and not torch.equal(a[0], draws[99][0])
    assert abs((a[0] * a[1]).mean().item()) < 5e-3                                # steps are uncorrelated


# ---------------------------------------------------------------------------------------------- conditioning
def test_time_film_and_cross_vec(ops, gen):
    from daclip_b200 import lib as L
    from oracle import unet_oracle as O
    B, F_ = 3, 1000
    sd = {"time_mlp.1.weight": rnd(gen, 256, 64, scale=0.2), "time_mlp.1.bias": rnd(gen, 256, scale=0.1),
          "time_mlp.3.weight": rnd(gen, 256, 256, scale=0.1), "time_mlp.3.bias": rnd(gen, 256, scale=0.1),
          "text_mlp.0.weight": rnd(gen, 256, 512, scale=0.05), "text_mlp.0.bias": rnd(gen, 256, scale=0.1),
          "text_mlp.2.weight": rnd(gen, 256, 256, scale=0.1), "text_mlp.2.bias": rnd(gen, 256, scale=0.1),
          "prompt": torch.rand(1, 256, generator=gen, device="cuda"),
          "prompt_mlp.weight": rnd(gen, 256, 256, scale=0.1), "prompt_mlp.bias": rnd(gen, 256, scale=0.1)}
    fw, fb = rnd(gen, F_, 256, scale=0.1), rnd(gen, F_, scale=0.1)
    ctx = rnd(gen, B, 512)
    ew = L.EmbedWeights()
    for name, key in [("time_w1", "time_mlp.1.weight"), ("time_b1", "time_mlp.1.bias"), ("time_w2", "time_mlp.3.weight"),
                      ("time_b2", "time_mlp.3.bias"), ("text_w1", "text_mlp.0.weight"), ("text_b1", "text_mlp.0.bias"),
                      ("text_w2", "text_mlp.2.weight"), ("text_b2", "text_mlp.2.bias"), ("prompt", "prompt"),
                      ("prompt_w", "prompt_mlp.weight"), ("prompt_b", "prompt_mlp.bias")]:
        setattr(ew, name, sd[key].data_ptr())
    ew.film_w, ew.film_b = fw.data_ptr(), fb.data_ptr()
    ew.nf, ew.time_dim, ew.ctx_dim, ew.F = 64, 256, 512, F_
    temb, film = torch.zeros(B, 256, device="cuda"), torch.zeros(B, F_, device="cuda")
    tdev = torch.zeros(1, device="cuda")
    pemb = torch.zeros(B, 256, device="cuda")
    ops.prompt_embed(ew, ctx, B, pemb)
    torch.cuda.synchronize()
    assert (pemb - O.prompt_embedding(sd, ctx)).abs().max().item() < 1e-4
    for t in (100.0, 37.0, 1.0):
        tdev.fill_(t)
        ops.time_film(ew, tdev, pemb, B, temb, film)
        torch.cuda.synchronize()
        te = O.time_embedding(sd, torch.tensor([t], device="cuda"), 64) + O.prompt_embedding(sd, ctx)
        ref = F.linear(F.silu(te), fw, fb)
        assert (film - ref).abs().max().item() < 2e-4 * max(1.0, ref.abs().max().item()), (film - ref).abs().max()
    w1, w2, b2 = rnd(gen, 256, 512, scale=0.05), rnd(gen, 256, 256, scale=0.1), rnd(gen, 256)
    y = torch.zeros(B, 256, device="cuda")
    ops.two_linear(ctx, w1, w2, b2, y)
    torch.cuda.synchronize()
    ref = F.linear(F.linear(ctx, w1), w2, b2)
    assert (y - ref).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item())


# ---------------------------------------------------------------------------------------------- ViT glue
def test_vit_glue(ops, gen):
    B, S, p, w, e = 3, 224, 32, 768, 512
    g = S // p
    img = rnd(gen, B, 3, S, S)
    patches = torch.zeros(B * g * g, 3 * p * p, device="cuda", dtype=torch.bfloat16)
    ops.vit_patchify(img, patches, B, S, p)
    torch.cuda.synchronize()
    ref = F.unfold(img, kernel_size=p, stride=p).transpose(1, 2).reshape(B * g * g, 3 * p * p)
    assert torch.equal(patches, bf(ref))
    L_ = g * g + 1
    pe = bf(rnd(gen, B, g * g, w))
    cls, pos = rnd(gen, w), rnd(gen, L_, w)
    lw, lb = 1 + 0.1 * rnd(gen, w), rnd(gen, w)
    tok = torch.zeros(B, L_, w, device="cuda", dtype=torch.float32)
    ops.vit_embed(pe, cls, pos, lw, lb, tok, B, L_, w)
    torch.cuda.synchronize()
    x = torch.cat([cls.expand(B, 1, w), pe.float()], 1) + pos
    assert (tok - F.layer_norm(x, (w,), lw, lb, 1e-5)).abs().max().item() < 1e-4
    proj = rnd(gen, w, e, scale=w ** -0.5)
    pooled = torch.zeros(B, e, device="cuda")
    ops.vit_pool(tok, B, L_, w, lw, lb, proj, pooled)
    torch.cuda.synchronize()
    ref = F.layer_norm(tok[:, 0], (w,), lw, lb, 1e-5) @ proj
    assert (pooled - ref).abs().max().item() < 1e-3
    text = rnd(gen, 10, e)
    logits, am = torch.zeros(B, 10, device="cuda"), torch.zeros(B, dtype=torch.int64, device="cuda")
    ops.degradation_argmax(pooled, text, logits, am)
    torch.cuda.synchronize()
    from oracle import daclip_oracle as D
    assert torch.equal(am, D.degradation_argmax(pooled, text))
    assert (logits - D.degradation_logits(pooled, text)).abs().max().item() < 1e-3
