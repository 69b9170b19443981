"""CPU tests: pin the oracle against outputs of the REFERENCE itself (tests/golden/*.pt, produced by
oracle/gen_golden.py from /root/reference with seeded synthetic weights).  fp32 on both sides; differences are
summation order only."""
import os

import pytest
import torch

from daclip_b200 import synthetic
from oracle import daclip_oracle as D
from oracle import sde_oracle as S
from oracle import unet_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "unet_sampler.pt"), weights_only=False)


@pytest.fixture(scope="module")
def net(gold):
    sd, kw = synthetic.unet_state_dict(gold["weights_seed"])
    assert kw == gold["ctor"]
    return sd, O.UNetConfig(**kw)


def test_schedule_known_answers(gold):
    s = S.Schedule(50, 100, "cosine", 0.005)
    for k in ("thetas", "sigmas", "thetas_cumsum", "sigma_bars"):
        assert torch.equal(getattr(s, k), gold["schedule"][k]), k
    assert torch.equal(s.dt, gold["schedule"]["dt"])
    # SURVEY.md section 8a row A1 known answers (probed from the reference)
    assert abs(float(s.dt) - 0.104093805) < 1e-8
    assert abs(float(s.thetas[1]) - 0.00169456005) < 1e-10
    assert abs(float(s.thetas[100]) - 0.999766588) < 1e-7
    assert abs(float(s.sigmas[100]) - 0.277264416) < 1e-7
    assert abs(float(s.sigma_bars[1]) - 0.00368262362) < 1e-9
    assert abs(float(s.sigma_bars[100]) - 0.196075976) < 1e-7
    assert abs(float(s.thetas_cumsum[100]) - 50.8994484) < 1e-4


def test_linear_and_constant_schedules():
    for name in ("linear", "constant"):
        s = S.Schedule(0.1, 20, name, 0.01)
        assert s.thetas.shape == (21,) and torch.isfinite(s.sigma_bars).all()
        assert abs(float(torch.exp(-s.thetas_cumsum[-1] * s.dt)) - 0.01) < 1e-6


def test_posterior_std_vanishes_at_t1():
    s = S.Schedule(50, 100, "cosine", 0.005)
    _, term2, std, _, _ = S.posterior_coeffs(s, 1)
    assert float(std) < 1e-10 and abs(float(term2) - 1.0) < 1e-6


def test_unet_forward_padded(gold, net):
    sd, cfg = net
    g = gold["fwd_pad"]
    B, H, W = g["shape"]
    inp = synthetic.restoration_inputs(B, H, W, T=1, seed=g["seed"])
    xt = inp["lq"] + inp["eps0"] * (50 / 255)
    with torch.no_grad():
        out = O.unet_forward(sd, cfg, xt, inp["lq"], g["time"], inp["text_context"], inp["image_context"])
    assert out.shape == g["out"].shape
    assert (out - g["out"]).abs().max().item() < 2e-4


def test_unet_forward_64(gold, net):
    sd, cfg = net
    g = gold["fwd_64"]
    inp = synthetic.restoration_inputs(1, 64, 64, T=1, seed=g["seed"])
    xt = inp["lq"] + inp["eps0"] * (50 / 255)
    with torch.no_grad():
        for t, key in ((100.0, "out_t100"), (1.0, "out_t1")):
            out = O.unet_forward(sd, cfg, xt, inp["lq"], t, inp["text_context"], inp["image_context"])
            assert (out - g[key]).abs().max().item() < 2e-4, key


@pytest.mark.parametrize("mode", ["sde", "posterior"])
def test_full_trajectory(gold, net, mode):
    sd, cfg = net
    g = gold["trajectory"]
    inp = synthetic.restoration_inputs(1, 32, 32, T=g["T"], seed=g["seed"])
    s = S.Schedule(50, g["T"], "cosine", 0.005)
    x_T = inp["lq"] + inp["eps0"] * s.max_sigma
    with torch.no_grad():
        x = S.reverse(s, O.make_denoiser(sd, cfg), x_T, inp["lq"], mode=mode, noise=inp["noise"],
                      text_context=inp["text_context"], image_context=inp["image_context"])
    assert (x - g[mode]).abs().max().item() < 1e-3


def test_cross_attention_single_token_is_broadcast(net):
    """attention.py:152-193 with a 1-token context == to_out(to_v(ctx)), independent of the query."""
    sd, cfg = net
    p = "mid_attn.fn.fn.transformer_blocks.0.attn2."
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 20, 512, generator=g)
    ctx = torch.randn(2, 1, 512, generator=g)
    full = O.attention(sd, p, x, ctx, heads=16)
    const = torch.nn.functional.linear(torch.nn.functional.linear(ctx, sd[p + "to_v.weight"]),
                                       sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])
    assert (full - const).abs().max().item() < 1e-5


def test_daclip_encode_and_argmax():
    g = torch.load(os.path.join(GOLD, "daclip.pt"), weights_only=False)
    sd = synthetic.daclip_visual_state_dict(g["weights_seed"])
    image = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(g["image_seed"]))
    taps = {}
    with torch.no_grad():
        img_f, deg_f = D.encode_image_control(sd, image, taps=taps)
    assert (img_f - g["image_features"]).abs().max().item() < 2e-3 * g["image_features"].abs().max().item()
    assert (deg_f - g["degra_features"]).abs().max().item() < 2e-3 * g["degra_features"].abs().max().item()
    assert (D.degradation_logits(deg_f, g["text_features"]) - g["logits"]).abs().max().item() < 5e-3
    assert torch.equal(D.degradation_argmax(deg_f, g["text_features"]), g["argmax"])
    assert torch.equal(D.degradation_argmax(g["degra_features"], g["text_features"]), g["argmax"])
    # the control signal must be alive (zero-init modules were randomised) and consumed in REVERSE order
    assert all(h.abs().max() > 1e-3 for h in taps["hiddens"])


def test_unet_wild_ir_variant_vs_reference_golden():
    """wild-ir ConditionalUNet (context_dim 768, no degradation prompt, scale 0.5 resampler pair): the oracle against
    outputs of the reference's wild-ir class (tests/golden/unet_wild.pt, oracle/gen_golden_wild.py)."""
    g = torch.load(os.path.join(GOLD, "unet_wild.pt"), weights_only=False)
    sd, kw = synthetic.unet_state_dict(g["weights_seed"], **g["ctor"])
    assert "downsample.weight" in sd and "upsample.1.bias" in sd and "prompt" not in sd
    cfg = O.UNetConfig(**kw)
    for case in g["cases"][:2]:
        B, H, W = case["shape"]
        inp = synthetic.restoration_inputs(B, H, W, T=1, seed=case["seed"], ctx_dim=768)
        xt = inp["lq"] + inp["eps0"] * (50 / 255)
        with torch.no_grad():
            out = O.unet_forward(sd, cfg, xt, inp["lq"], case["time"], inp["text_context"], inp["image_context"])
        assert out.shape == case["out"].shape
        assert (out - case["out"]).abs().max().item() < 2e-4


def test_unet_without_image_context_vs_reference_golden():
    """The class-default construction (use_image_context=False, DenoisingUNet_arch.py:22-23): LinearAttention at every
    level incl. 512 channels; the oracle against the reference's own outputs (tests/golden/unet_noctx.pt)."""
    g = torch.load(os.path.join(GOLD, "unet_noctx.pt"), weights_only=False)
    sd, kw = synthetic.unet_state_dict(g["weights_seed"], **g["ctor"])
    assert "mid_attn.fn.fn.to_qkv.weight" in sd and not any("transformer_blocks" in k for k in sd)
    cfg = O.UNetConfig(**kw)
    for case in g["cases"]:
        B, H, W = case["shape"]
        inp = synthetic.restoration_inputs(B, H, W, T=1, seed=case["seed"])
        xt = inp["lq"] + inp["eps0"] * (50 / 255)
        with torch.no_grad():
            out = O.unet_forward(sd, cfg, xt, inp["lq"], case["time"], inp["text_context"], None)
        assert out.shape == case["out"].shape
        assert (out - case["out"]).abs().max().item() < 2e-4


def test_daclip_vit_l14_encode_vs_reference_golden():
    """wild-ir encoder (daclip_ViT-L-14: patch 14, 257 tokens, width 1024, 24 layers, 16 heads, embed 768): oracle vs
    the reference model's own features (tests/golden/daclip_l14.pt)."""
    g = torch.load(os.path.join(GOLD, "daclip_l14.pt"), weights_only=False)
    sd = synthetic.daclip_visual_state_dict(g["weights_seed"], arch="ViT-L-14")
    image = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(g["image_seed"]))
    cfg = D.ViTConfig(image_size=224, patch=14, width=1024, layers=24, heads=16, embed_dim=768)
    with torch.no_grad():
        img_f, deg_f = D.encode_image_control(sd, image, cfg)
    assert img_f.shape == (2, 768)
    assert (img_f - g["image_features"]).abs().max().item() < 2e-3 * g["image_features"].abs().max().item()
    assert (deg_f - g["degra_features"]).abs().max().item() < 2e-3 * g["degra_features"].abs().max().item()


@pytest.mark.parametrize("mode", ["sde", "posterior"])
def test_reduced_step_sampling_vs_reference_golden(net, mode):
    """IRSDE(T=100, sample_T=20) (sde_utils.py:87-89,195-202): tables for 20 steps, the network queried at 5 t."""
    sd, cfg = net
    g = torch.load(os.path.join(GOLD, "sampler_reduced.pt"), weights_only=False)
    ST = g["sample_T"]
    inp = synthetic.restoration_inputs(1, 32, 32, T=ST, seed=g["seed"])
    s = S.Schedule(50, ST, "cosine", 0.005)
    assert torch.equal(s.dt, g["dt"]) and torch.equal(s.sigma_bars, g["sigma_bars"])
    x_T = inp["lq"] + inp["eps0"] * s.max_sigma
    with torch.no_grad():
        x = S.reverse(s, O.make_denoiser(sd, cfg), x_T, inp["lq"], mode=mode, noise=inp["noise"],
                      time_scale=g["T"] / ST, text_context=inp["text_context"], image_context=inp["image_context"])
    assert (x - g[mode]).abs().max().item() < 1e-3


@pytest.mark.parametrize("arch,heads", [("ViT-B-32", 8), ("ViT-L-14", 12)])
def test_daclip_encode_text_vs_reference_golden(arch, heads):
    """the oracle's encode_text against CLIP.encode_text of the reference (tests/golden/daclip_text.pt:
    degradation prompts, two captions, a one-letter prompt and a truncated 60-word prompt)."""
    g = torch.load(os.path.join(GOLD, "daclip_text.pt"), weights_only=False)
    sd = synthetic.daclip_text_state_dict(g[arch]["weights_seed"], arch=arch)
    with torch.no_grad():
        f = D.encode_text(sd, g["tokens"].long(), heads=heads)
    ref = g[arch]["features"]
    assert f.shape == ref.shape
    assert (f - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()


def test_tokenizer_vs_reference_golden_and_live_reference():
    """daclip_b200.tokenizer against the token ids the reference tokenizer produced (golden), and - where the reference
    checkout and its vocabulary are present (this container, not the GPU box) - against the reference tokenizer live on
    strings with unicode, html entities, contractions, special tokens and over-long input."""
    ref_root = "/root/reference/universal-image-restoration"
    vocab = os.path.join(ref_root, "open_clip", "bpe_simple_vocab_16e6.txt.gz")
    if not os.path.isfile(vocab):
        pytest.skip("the BPE merge table is a reference asset that does not travel")
    from daclip_b200.tokenizer import SimpleTokenizer
    tk = SimpleTokenizer(vocab)
    g = torch.load(os.path.join(GOLD, "daclip_text.pt"), weights_only=False)
    assert torch.equal(tk(g["prompts"]).to(torch.int32), g["tokens"])
    assert tk.vocab_size == 49408 and tk.eot_token == 49407
    import sys
    import types
    sys.path.insert(0, ref_root)
    had = "ftfy" in sys.modules
    sys.modules.setdefault("ftfy", types.SimpleNamespace(fix_text=lambda s: s))
    try:
        from open_clip import tokenizer as RT
    finally:
        sys.path.remove(ref_root)
        if not had:
            sys.modules.pop("ftfy", None)
    texts = ["A photo of a cat's   whiskers, isn't it?  ", "Ünïcödé çafé — naïve 12345 résumé",
             "&amp;lt;b&amp;gt; html &quot;entities&quot;", "emoji 😀 and 中文 текст",
             "<start_of_text> weird <end_of_text> specials", "x" * 300, "", "I'LL we'RE THEY'd 's"]
    assert torch.equal(tk(texts), RT.tokenize(texts))
    ids = tk.encode(texts[0])
    assert tk.decode(ids) == RT._tokenizer.decode(ids)
