"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the drop-in modules
keep the reference's state-dict layout, weight repacking is exact, and the product path refuses to run on CPU."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from daclip_b200 import lib
    l = lib.load()
    header = open(os.path.join(ROOT, "include", "dac_b200.h")).read()
    declared = set(re.findall(r"\b(dac_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(lib.SYMBOLS), declared ^ set(lib.SYMBOLS)
    for name in declared:
        assert hasattr(l, name), name
    assert l.dac_version() >= 100
    a, b = ctypes.c_int32(), ctypes.c_int32()
    l.dac_abi_sizes(ctypes.byref(a), ctypes.byref(b))
    assert a.value == ctypes.sizeof(lib.ConvDesc) and b.value == ctypes.sizeof(lib.EmbedWeights)


def test_no_cpu_fallback():
    from daclip_b200 import lib
    from daclip_b200.unet import ConditionalUNet
    from daclip_b200.sde import IRSDE
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    m = ConditionalUNet(3, 3, 64, [1, 2, 4, 8], 512, True, True)
    x = torch.zeros(1, 3, 32, 32)
    with pytest.raises(lib.DacError):
        m(x, x, 10.0, text_context=torch.zeros(1, 512), image_context=torch.zeros(1, 512))
    sde = IRSDE(50, T=100, schedule="cosine", eps=0.005, device="cpu")
    sde.set_mu(x)
    with pytest.raises(lib.DacError):
        sde.reverse_posterior_step(x, x, 5)


def test_state_dict_layout_matches_golden_ctor():
    from daclip_b200 import synthetic
    from daclip_b200.unet import ConditionalUNet
    sd, kw = synthetic.unet_state_dict(0)
    m = ConditionalUNet(**kw)
    assert len(sd) == 224 and sum(v.numel() for v in sd.values()) == 48975747
    m.load_state_dict(sd, strict=True)
    # a few keys/shapes called out in SURVEY.md section 8b
    assert tuple(sd["downs.0.0.block1.proj.weight"].shape) == (64, 64, 3, 3)
    assert tuple(sd["downs.3.2.fn.fn.transformer_blocks.0.attn2.to_k.weight"].shape) == (256, 512)
    assert tuple(sd["ups.0.3.1.weight"].shape) == (256, 512, 3, 3)
    assert tuple(sd["prompt"].shape) == (1, 256)
    # 'module.' prefix of DataParallel checkpoints (base_model.py:92-105) round-trips
    wrapped = torch.nn.DataParallel(m) if False else None  # construction needs no GPU; wrapping is plumbing
    del wrapped


def _emulate(pw, x_nchw, B, H, W):
    """Executes a PackedWeight the way the kernel does (taps, groups, stride, output scatter) with torch ops."""
    cin = pw.w.shape[-1]
    s = pw.stride
    OH, OW = (H // s, W // s) if s == 2 else (H, W)
    out = torch.zeros(B, pw.cout, OH * pw.out_scale, OW * pw.out_scale)
    pad = 4
    xp = F.pad(x_nchw, (pad, pad, pad, pad))
    ntaps = len(pw.taps[0])
    for g in range(pw.ngroups):
        acc = torch.zeros(B, pw.cout, OH, OW)
        for t, (dy, dx) in enumerate(pw.taps[g]):
            wt = pw.w[g * ntaps + t, :pw.cout].float()                    # [cout, cin]
            ys = torch.arange(OH) * s + dy + pad
            xs = torch.arange(OW) * s + dx + pad
            patch = xp[:, :, ys][:, :, :, xs]                              # [B, cin, OH, OW]
            acc += torch.einsum("oc,bchw->bohw", wt, patch)
        oy, ox = pw.out_off[g]
        out[:, :, oy::pw.out_scale, ox::pw.out_scale] = acc
    return out


def test_weight_packing_semantics():
    from daclip_b200 import ops
    g = torch.Generator().manual_seed(0)
    bfr = lambda t: t.to(torch.bfloat16).float()
    B, H, W, cin, cout = 2, 8, 12, 64, 32
    x = bfr(torch.randn(B, cin, H, W, generator=g))
    w3 = bfr(torch.randn(cout, cin, 3, 3, generator=g) * 0.05)
    assert torch.allclose(_emulate(ops.pack_conv(w3), x, B, H, W), F.conv2d(x, w3, padding=1), atol=1e-4)
    w4 = bfr(torch.randn(cout, cin, 4, 4, generator=g) * 0.05)
    assert torch.allclose(_emulate(ops.pack_conv(w4, stride=2, pad=1), x, B, H, W),
                          F.conv2d(x, w4, stride=2, padding=1), atol=1e-4)
    wu = torch.randn(cout, cin, 3, 3, generator=g) * 0.05
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), wu, padding=1)
    assert torch.allclose(_emulate(ops.pack_upsample_conv(wu), x, B, H, W), ref, atol=2e-2)
    # stem: 7x7 over 6 channels == 7 vertical taps over the (kx, c)-packed 64 channels
    ws = bfr(torch.randn(64, 6, 7, 7, generator=g) * 0.05)
    x6 = bfr(torch.randn(B, 6, H, W, generator=g))
    packed = torch.zeros(B, 64, H, W)
    xp = F.pad(x6, (3, 3, 0, 0))
    for kx in range(7):
        packed[:, kx * 8:kx * 8 + 6] = xp[:, :, :, kx:kx + W]
    assert torch.allclose(_emulate(ops.pack_stem(ws), packed, B, H, W), F.conv2d(x6, ws, padding=3), atol=1e-4)
    # ... and in pixel-pair form: window kx' = 0..7 around a pair, 128 weight rows, output [.., W/2, 128] == [.., W, 64]
    packed2 = torch.zeros(B, 64, H, W // 2)
    xp2 = F.pad(x6, (3, 4, 0, 0))
    for kx in range(8):
        packed2[:, kx * 8:kx * 8 + 6] = xp2[:, :, :, kx:kx + W:2]
    got2 = _emulate(ops.pack_stem_pair(ws), packed2, B, H, W // 2)               # [B, 128, H, W/2]
    got2 = got2.reshape(B, 2, 64, H, W // 2).permute(0, 2, 3, 4, 1).reshape(B, 64, H, W)
    assert torch.allclose(got2, F.conv2d(x6, ws, padding=3), atol=1e-4)
    # GEGLU interleave
    wg, bg = torch.randn(1024, 64, generator=g) * 0.125, torch.randn(1024, generator=g)
    pw, bperm = ops.pack_geglu(wg, bg, block_n=256)
    y = F.linear(torch.randn(5, 64, generator=torch.Generator().manual_seed(1)), wg, bg)
    yp = F.linear(torch.randn(5, 64, generator=torch.Generator().manual_seed(1)), pw.w[0].float(), bperm)
    val, gate = y.chunk(2, -1)
    tiles = yp.reshape(5, 4, 256)
    assert torch.allclose(tiles[:, :, :128].reshape(5, 512), val, atol=0.05)
    assert torch.allclose(tiles[:, :, 128:].reshape(5, 512), gate, atol=0.05)


def test_tile_and_block_choices():
    from daclip_b200 import ops
    for oh, ow in [(256, 256), (128, 128), (64, 64), (32, 32), (6, 6), (1, 800), (48, 80)]:
        th, tw = ops.choose_tile(oh, ow)
        assert th * tw == 128 and tw % 8 == 0
    assert ops.choose_block_n(64) == (64, 64) and ops.choose_block_n(3) == (16, 16)
    assert ops.choose_block_n(384) == (128, 384) and ops.choose_block_n(512) == (256, 512)
    assert ops.choose_block_n(2304) == (256, 2304)


def test_irsde_host_tables_match_oracle():
    from daclip_b200.sde import IRSDE
    from oracle import sde_oracle as S
    sde = IRSDE(50, T=100, schedule="cosine", eps=0.005, device="cpu")
    s = S.Schedule(50, 100, "cosine", 0.005)
    for k in ("thetas", "sigmas", "thetas_cumsum", "sigma_bars"):
        assert torch.equal(getattr(sde, k), getattr(s, k))
    assert torch.equal(sde.dt, s.dt)
    for t in (100, 50, 1):
        for a, b in zip(sde._posterior_coef(t), S.posterior_coeffs(s, t)):
            assert float(a) == float(b)
    assert sde.sample_scale == 1.0 and abs(sde.max_sigma - 50 / 255) < 1e-12


def test_pixel_pair_packing_semantics():
    """The pixel-pair schedule of the conv kernel (dac_conv_desc.pair), emulated on the CPU from the packed weight block:
    on the [B, H, W/2, 2C] view, per (64-channel slice, ky) the even chunk multiplies rows [64,192) of the block at pair i
    and rows [0,64) at pair i+1 (into the odd half), the odd chunk rows [0,128) at pair i and rows [128,192) at pair i-1
    (into the even half) - together exactly the 3x3 convolution, for one source and for a two-source concat."""
    from daclip_b200 import ops
    g = torch.Generator().manual_seed(5)
    bfr = lambda t: t.to(torch.bfloat16).float()
    B, H, W = 2, 6, 10
    for cin in (64, 128):
        x = bfr(torch.randn(B, cin, H, W, generator=g))
        w = bfr(torch.randn(64, cin, 3, 3, generator=g) * 0.05)
        wall = ops.pack_conv_pair(w).float()                    # [9][192][cin]: 1-CTA layout + the two per-rank layouts
        assert wall.shape == (9, 192, cin)
        wp = wall[:3]                                           # [3][192][cin]
        xp = F.pad(x.permute(0, 2, 3, 1), (0, 0, 2, 2, 1, 1))  # NHWC, two pixels = one pair of zero padding left and right
        out = torch.zeros(B, H, W // 2, 128)
        for s in range(cin // 64):
            sl = slice(64 * s, 64 * s + 64)
            for ky in range(3):
                blk = wp[ky][:, sl]                             # [192][64]
                rows = xp[:, ky:ky + H]                         # input row y + ky - 1
                even = lambda d: rows[:, :, 2 + 2 * d:2 + 2 * d + W:2, sl]    # x[2(i + d)]
                odd = lambda d: rows[:, :, 3 + 2 * d:3 + 2 * d + W:2, sl]     # x[2(i + d) + 1]
                out += even(0) @ blk[64:192].T                  # N = 128: centre of 2i | left neighbour of 2i+1
                out += odd(0) @ blk[0:128].T                    # N = 128: right neighbour of 2i | centre of 2i+1
                out[..., :64] += odd(-1) @ blk[128:192].T       # N = 64: left neighbour of pixel 2i
                out[..., 64:] += even(1) @ blk[0:64].T          # N = 64: right neighbour of pixel 2i+1
        got = out.reshape(B, H, W, 64).permute(0, 3, 1, 2)      # the wide output IS the NHWC output
        assert torch.allclose(got, F.conv2d(x, w, padding=1), atol=1e-4)
        # CTA-pair build (tcgen05.mma.cta_group::2): CTA r supplies the B rows of output columns [N/2 r, +N/2) from ITS block
        # [E; O; S0; S2] (blocks 3 + 3 r + ky): N = 128 windows = [E_0; E_1] / [O_0; O_1], N = 64 windows = [S0_0; S0_1] /
        # [S2_0; S2_1] - the same four products as above
        out2 = torch.zeros(B, H, W // 2, 128)
        for sidx in range(cin // 64):
            sl = slice(64 * sidx, 64 * sidx + 64)
            for ky in range(3):
                r0, r1 = wall[3 + ky][:, sl], wall[6 + ky][:, sl]
                rows = xp[:, ky:ky + H]
                even = lambda d: rows[:, :, 2 + 2 * d:2 + 2 * d + W:2, sl]
                odd = lambda d: rows[:, :, 3 + 2 * d:3 + 2 * d + W:2, sl]
                out2 += even(0) @ torch.cat([r0[0:64], r1[0:64]]).T          # E: centre window of the even chunk
                out2 += odd(0) @ torch.cat([r0[64:128], r1[64:128]]).T       # O: centre window of the odd chunk
                out2[..., :64] += odd(-1) @ torch.cat([r0[128:160], r1[128:160]]).T   # S0 = W(kx=0): left neighbour of pixel 2i
                out2[..., 64:] += even(1) @ torch.cat([r0[160:192], r1[160:192]]).T   # S2 = W(kx=2): right neighbour of 2i+1
        assert torch.equal(out2, out)
    w3 = torch.randn(3, 64, 3, 3, generator=g)
    wp = ops.pack_conv_pair(F.pad(w3, (0, 0, 0, 0, 0, 0, 0, 13)))
    assert wp.shape == (3, 48, 64) and torch.equal(wp[1, 16:19].float(), bfr(w3[:, :, 1, 1]))


def test_context_partial_slots_cover_every_image():
    """dac_linattn_ctx_slots (no GPU needed: 148 SMs assumed without a device) against a brute-force walk of the kernels'
    tile ranges: CTA b owns tiles [T b / G, T (b+1) / G); the slot count must cover the widest span of CTAs over any
    image, for one record per CTA (k|v kernel) and two (KVCTX epilogue groups)."""
    from daclip_b200 import lib as L
    lib = L.load()
    for B, tpi in [(16, 512), (16, 32), (1, 512), (1, 3), (5, 37), (40, 1), (16, 18), (3, 2048), (128, 8)]:
        T = B * tpi
        G = min(T, 148)
        owner = []
        for b in range(G):
            owner += [b] * (T * (b + 1) // G - T * b // G)
        assert len(owner) == T
        span = max(owner[(i + 1) * tpi - 1] - owner[i * tpi] + 1 for i in range(B))
        assert lib.dac_linattn_ctx_slots(B, tpi, 1) == span, (B, tpi)
        assert lib.dac_linattn_ctx_slots(B, tpi, 2) == 2 * span
    assert lib.dac_linattn_ctx_slots(0, 4, 1) == 0


def test_option_yaml_contract(tmp_path):
    """options.parse + dict_to_nonedict on the reference's inference option file (tests/golden/options_test.yml, copied
    byte for byte by oracle/gen_golden_options.py) against what the REFERENCE's own options.parse / dict_to_nonedict
    produce (tests/golden/options_test.json; config/daclip-sde/options.py:18-120)."""
    import json
    from daclip_b200 import options
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    want = json.load(open(os.path.join(gold, "options_test.json")))
    before = os.environ.get("CUDA_VISIBLE_DEVICES")
    opt = options.dict_to_nonedict(options.parse(os.path.join(gold, "options_test.yml"), is_train=False, root=str(tmp_path)))
    assert os.environ.get("CUDA_VISIBLE_DEVICES") == before            # the launcher owns it unless asked
    got = json.loads(json.dumps(opt))
    root = got["path"].pop("root")
    assert root == str(tmp_path)
    assert got["path"].pop("results_root") == os.path.join(root, "results", "daclip-sde", "universal-ir")
    assert got["path"].pop("log") == os.path.join(root, "results", "daclip-sde", "universal-ir")
    assert got == want["opt"]
    assert isinstance(opt, options.NoneDict) and isinstance(opt["network_G"]["setting"], options.NoneDict)
    assert opt["suffix"] is None and opt["crop_border"] is None and opt["no_such_key"] is None
    assert opt["path"]["strict_load"] is None and opt["datasets"]["test1"]["no_such_key"] is None
    assert want["missing_keys_read_as"] == {"suffix": None, "crop_border": None, "no_such_key": None, "path.strict_load": None}
    # what the restoration script reads from it (test.py:68-85)
    assert opt["sde"] == {"max_sigma": 50, "T": 100, "schedule": "cosine", "eps": 0.005, "sampling_mode": "posterior"}
    assert opt["network_G"]["which_model_G"] == "ConditionalUNet" and opt["degradation"]["scale"] == 4
