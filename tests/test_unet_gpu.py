"""GPU parity of the denoiser and of the full reverse-time loops, through the drop-in Python API
(daclip_b200.ConditionalUNet / IRSDE -> C ABI -> sm_100a kernels).

Checkers: (1) committed outputs of the REFERENCE modules (tests/golden/unet_sampler.pt, fp32, CPU);
(2) the fp32 oracle evaluated on the same seeded inputs at sizes the goldens do not cover.
Tolerances (north_star): final restored image max-abs <= 2e-2 and PSNR >= 45 dB (images in [0,1]) for the
bf16 tensor-core path; single denoiser evaluations within 1 % of the output range (measured on a B200: 3.8e-3 ... 4.4e-3
for every case below, 1.5e-2 for the worst single layer - the bounds are ~2x those; gpurun_out/unet_errors.json).
"""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "unet_sampler.pt"), weights_only=False)


@pytest.fixture(scope="module")
def model(cuda, gold):
    from daclip_b200 import synthetic
    from daclip_b200.unet import ConditionalUNet
    sd, kw = synthetic.unet_state_dict(gold["weights_seed"])
    m = ConditionalUNet(**kw)
    m.load_state_dict(sd, strict=True)
    return m.to(cuda).eval(), sd, kw


def psnr(a, b):
    mse = torch.mean((a.clamp(0, 1) - b.clamp(0, 1)) ** 2).item()
    return 99.0 if mse == 0 else 10 * math.log10(1.0 / mse)


_ACHIEVED = {}


def rel_err(got, ref):
    """max |got - ref| / max |ref|; every value is also logged per calling test (gpurun_out/unet_errors.json), which is
    where the bounds below come from: 2x the worst value measured on a B200."""
    import inspect
    import json
    e = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-6)
    who = inspect.stack()[1].function
    _ACHIEVED[who] = max(_ACHIEVED.get(who, 0.0), e)
    out = os.path.join(os.path.dirname(GOLD), "..", "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "unet_errors.json"), "w") as f:
            json.dump(_ACHIEVED, f, indent=1, sort_keys=True)
    return e


def test_forward_vs_reference_golden_padded(model, gold):
    from daclip_b200 import synthetic
    m, _, _ = model
    g = gold["fwd_pad"]
    B, H, W = g["shape"]
    inp = synthetic.restoration_inputs(B, H, W, T=1, seed=g["seed"])
    xt = (inp["lq"] + inp["eps0"] * (50 / 255)).cuda()
    out = m(xt, inp["lq"].cuda(), g["time"], text_context=inp["text_context"].cuda(),
            image_context=inp["image_context"].cuda())
    assert out.shape == g["out"].shape and out.dtype == torch.float32
    e = rel_err(out.cpu(), g["out"])
    assert e < 1e-2, f"relative max error {e:.4f}"


def test_forward_vs_reference_golden_64(model, gold):
    from daclip_b200 import synthetic
    m, _, _ = model
    g = gold["fwd_64"]
    inp = synthetic.restoration_inputs(1, 64, 64, T=1, seed=g["seed"])
    xt = (inp["lq"] + inp["eps0"] * (50 / 255)).cuda()
    for t, key in ((100.0, "out_t100"), (1.0, "out_t1")):
        out = m(xt, inp["lq"].cuda(), t, text_context=inp["text_context"].cuda(),
                image_context=inp["image_context"].cuda())
        e = rel_err(out.cpu(), g[key])
        assert e < 1e-2, f"{key}: relative max error {e:.4f}"


def test_forward_per_layer_vs_oracle(model):
    """Every block output of one 64x64 evaluation against the oracle's taps: localises a broken layer."""
    from daclip_b200 import synthetic
    from oracle import unet_oracle as O
    m, sd, kw = model
    cfg = O.UNetConfig(**kw)
    inp = {k: v.cuda() for k, v in synthetic.restoration_inputs(2, 64, 64, T=1, seed=5).items()}
    sdc = {k: v.cuda() for k, v in sd.items()}
    xt = inp["lq"] + inp["eps0"] * (50 / 255)
    taps = {}
    with torch.no_grad():
        ref = O.unet_forward(sdc, cfg, xt, inp["lq"], 63.0, inp["text_context"], inp["image_context"], taps=taps)
    out = m(xt, inp["lq"], 63.0, text_context=inp["text_context"], image_context=inp["image_context"])
    eng = m.engine(2, 64, 64)
    worst = []
    for name, buf in eng.taps.items():
        got = buf.float().permute(0, 3, 1, 2)
        e = rel_err(got, taps[name])
        worst.append((e, name))
        assert e < 3e-2, f"{name}: relative max error {e:.4f}"      # worst layer measured: 1.5e-2
    assert rel_err(out, ref) < 1e-2, sorted(worst)[-3:]


def test_forward_256_vs_oracle(model):
    from daclip_b200 import synthetic
    from oracle import unet_oracle as O
    m, sd, kw = model
    cfg = O.UNetConfig(**kw)
    inp = {k: v.cuda() for k, v in synthetic.restoration_inputs(2, 256, 256, T=1, seed=6).items()}
    sdc = {k: v.cuda() for k, v in sd.items()}
    xt = inp["lq"] + inp["eps0"] * (50 / 255)
    with torch.no_grad():
        ref = O.unet_forward(sdc, cfg, xt, inp["lq"], 80.0, inp["text_context"], inp["image_context"])
    out = m(xt, inp["lq"], 80.0, text_context=inp["text_context"], image_context=inp["image_context"])
    assert rel_err(out, ref) < 1e-2, rel_err(out, ref)
    # a second call with different inputs must not reuse stale state (CUDA graph replays static buffers)
    out2 = m(xt * 0.5, inp["lq"], 3.0, text_context=inp["text_context"], image_context=inp["image_context"])
    with torch.no_grad():
        ref2 = O.unet_forward(sdc, cfg, xt * 0.5, inp["lq"], 3.0, inp["text_context"], inp["image_context"])
    assert rel_err(out2, ref2) < 1e-2


@pytest.mark.parametrize("mode", ["sde", "posterior"])
def test_full_trajectory_vs_reference_golden(model, gold, cuda, mode):
    """T=100 reverse loop with injected noise vs the reference's own loop (fp32, CPU)."""
    from daclip_b200 import synthetic
    from daclip_b200.sde import IRSDE
    m, _, _ = model
    g = gold["trajectory"]
    inp = {k: v.cuda() for k, v in synthetic.restoration_inputs(1, 32, 32, T=g["T"], seed=g["seed"]).items()}
    sde = IRSDE(max_sigma=50, T=g["T"], schedule="cosine", eps=0.005, device=cuda)
    sde.set_model(torch.nn.DataParallel(m, device_ids=[0]))       # the wrapper the reference applies
    sde.set_mu(inp["lq"])
    x_T = inp["lq"] + inp["eps0"] * sde.max_sigma
    fn = sde.reverse_sde if mode == "sde" else sde.reverse_posterior
    x = fn(x_T, noise=inp["noise"], text_context=inp["text_context"], image_context=inp["image_context"]).cpu()
    ref = g[mode]
    err = (x - ref).abs().max().item()
    assert err <= 2e-2, f"max-abs {err:.4g}"
    assert psnr(x, ref) >= 45.0, psnr(x, ref)
    # the generic (non-fused) loop - any callable as the model - must agree with the fused one
    sde.set_model(lambda x_, mu, t, **kw: m(x_, mu, t, **kw))
    x2 = fn(x_T, noise=inp["noise"], text_context=inp["text_context"], image_context=inp["image_context"]).cpu()
    # (same kernels, same order: the eager per-step path and the fused loop agree bit for bit - no atomics anywhere)
    assert torch.equal(x2, x), (x2 - x).abs().max().item()


def test_trajectory_256_batch_vs_oracle(model, cuda):
    """Batch 2 at 256x256, T=8 posterior steps taken from the top of the schedule, vs the oracle on the GPU."""
    from daclip_b200 import synthetic
    from daclip_b200.sde import IRSDE
    from oracle import sde_oracle as S
    from oracle import unet_oracle as O
    m, sd, kw = model
    cfg = O.UNetConfig(**kw)
    T = 100
    inp = {k: v.cuda() for k, v in synthetic.restoration_inputs(2, 256, 256, T=8, seed=8).items()}
    sde = IRSDE(max_sigma=50, T=T, schedule="cosine", eps=0.005, device=cuda)
    sde.set_model(m)
    sde.set_mu(inp["lq"])
    x_T = inp["lq"] + inp["eps0"] * sde.max_sigma
    sdc = {k: v.cuda() for k, v in sd.items()}
    den = O.make_denoiser(sdc, cfg)
    sch = S.Schedule(50, T, "cosine", 0.005)        # 0-dim CPU coefficients broadcast onto CUDA tensors
    x_ref, x = x_T.clone(), x_T.clone()
    with torch.no_grad():
        for i, t in enumerate(range(T, T - 8, -1)):
            n = den(x_ref, inp["lq"], float(t), text_context=inp["text_context"], image_context=inp["image_context"])
            x_ref = S.posterior_step(sch, x_ref, inp["lq"], n, inp["noise"][i], t)
            net = m(x, inp["lq"], float(t), text_context=inp["text_context"], image_context=inp["image_context"])
            x = sde.reverse_posterior_step(x, net, t, eps=inp["noise"][i])
    assert (x - x_ref).abs().max().item() <= 2e-2
    assert psnr(x, x_ref) >= 45.0


def test_forward_512_vs_oracle(model):
    """BASELINE config C4 shape (512x512): 4096-token self-attention, larger conv grids."""
    from daclip_b200 import synthetic
    from oracle import unet_oracle as O
    m, sd, kw = model
    cfg = O.UNetConfig(**kw)
    inp = {k: v.cuda() for k, v in synthetic.restoration_inputs(1, 512, 512, T=1, seed=9).items()}
    sdc = {k: v.cuda() for k, v in sd.items()}
    xt = inp["lq"] + inp["eps0"] * (50 / 255)
    with torch.no_grad():
        ref = O.unet_forward(sdc, cfg, xt, inp["lq"], 42.0, inp["text_context"], inp["image_context"])
    out = m(xt, inp["lq"], 42.0, text_context=inp["text_context"], image_context=inp["image_context"])
    assert rel_err(out, ref) < 1e-2, rel_err(out, ref)


def test_odd_sizes_and_batch(model):
    """Reflect padding in both directions, batch 3, non-square (reference: arch.py:111-116,172)."""
    from daclip_b200 import synthetic
    from oracle import unet_oracle as O
    m, sd, kw = model
    cfg = O.UNetConfig(**kw)
    sdc = {k: v.cuda() for k, v in sd.items()}
    for (B, H, W) in [(3, 50, 70), (1, 17, 33)]:
        inp = {k: v.cuda() for k, v in synthetic.restoration_inputs(B, H, W, T=1, seed=12).items()}
        xt = inp["lq"] + inp["eps0"] * (50 / 255)
        with torch.no_grad():
            ref = O.unet_forward(sdc, cfg, xt, inp["lq"], 7.0, inp["text_context"], inp["image_context"])
        out = m(xt, inp["lq"], 7.0, text_context=inp["text_context"], image_context=inp["image_context"])
        assert out.shape == (B, 3, H, W)
        assert rel_err(out, ref) < 1e-2, (B, H, W, rel_err(out, ref))


def test_wrapper_test_api(model, cuda):
    """DenoisingModel-style wrapper: feed_data / test(sde, mode) / get_current_visuals (denoising_model.py:121-173)."""
    from daclip_b200 import synthetic
    from daclip_b200.model import DenoisingModel
    from daclip_b200.sde import IRSDE
    m, sd, kw = model
    opt = {"gpu_ids": [0], "is_train": False, "dist": False,
           "network_G": {"which_model_G": "ConditionalUNet", "setting": dict(kw)},
           "path": {"pretrain_model_G": None, "strict_load": True}, "train": None}
    wrapper = DenoisingModel(opt)
    wrapper.load_state_dict_into_model(sd)
    sde = IRSDE(max_sigma=50, T=100, sample_T=-1, schedule="cosine", eps=0.005, device=wrapper.device)
    sde.set_model(wrapper.model)
    inp = synthetic.restoration_inputs(1, 32, 32, T=100, seed=3)
    lq = inp["lq"]
    torch.manual_seed(0)
    noisy = sde.noise_state(lq)                                # CPU in, CPU out like the reference call site
    assert noisy.device.type == "cpu" and noisy.shape == lq.shape
    wrapper.feed_data(noisy, lq, lq, text_context=inp["text_context"].cuda(), image_context=inp["image_context"].cuda())
    for mode in ("posterior", "sde"):
        wrapper.test(sde, mode=mode)
        vis = wrapper.get_current_visuals()
        assert set(vis) == {"Input", "Output", "GT"} and vis["Output"].shape == (3, 32, 32)
        assert torch.isfinite(vis["Output"]).all() and vis["Output"].device.type == "cpu"


def test_wild_ir_variant_vs_reference_golden(cuda):
    """the wild-ir denoiser (context_dim 768, use_degra_context false, scale 0.5 = extra Downsample /
    Upsample pair around the UNet) against outputs of the reference's wild-ir class, incl. a reflect-padded size."""
    from daclip_b200 import synthetic
    from daclip_b200.unet import ConditionalUNet
    g = torch.load(os.path.join(GOLD, "unet_wild.pt"), weights_only=False)
    sd, kw = synthetic.unet_state_dict(g["weights_seed"], **g["ctor"])
    m = ConditionalUNet(**kw)
    m.load_state_dict(sd, strict=True)
    m = m.to(cuda).eval()
    for case in g["cases"]:
        B, H, W = case["shape"]
        inp = {k: v.cuda() for k, v in synthetic.restoration_inputs(B, H, W, T=1, seed=case["seed"], ctx_dim=768).items()}
        xt = inp["lq"] + inp["eps0"] * (50 / 255)
        out = m(xt, inp["lq"], case["time"], text_context=inp["text_context"], image_context=inp["image_context"])
        ref = case["out"].cuda()
        assert out.shape == ref.shape
        assert rel_err(out, ref) < 1e-2, (case["shape"], rel_err(out, ref))


def test_no_image_context_variant_vs_reference_golden(cuda):
    """ConditionalUNet(use_image_context=False) - the reference class default (DenoisingUNet_arch.py:22-23,84-85,105): every
    level is a LinearAttention, the 512-channel ones (mid, ups.0) through the plain to_out epilogue + LayerNorm/residual
    pass - against the reference's outputs, and IRSDE.reverse_ode (sde_utils.py:282-294: forwards no contexts, so it only
    runs with this construction) over the full T = 100 against the reference's own ODE loop."""
    from daclip_b200 import synthetic
    from daclip_b200.sde import IRSDE
    from daclip_b200.unet import ConditionalUNet
    g = torch.load(os.path.join(GOLD, "unet_noctx.pt"), weights_only=False)
    sd, kw = synthetic.unet_state_dict(g["weights_seed"], **g["ctor"])
    m = ConditionalUNet(**kw)
    m.load_state_dict(sd, strict=True)
    m = m.to(cuda).eval()
    for case in g["cases"]:
        B, H, W = case["shape"]
        inp = {k: v.cuda() for k, v in synthetic.restoration_inputs(B, H, W, T=1, seed=case["seed"]).items()}
        xt = inp["lq"] + inp["eps0"] * (50 / 255)
        out = m(xt, inp["lq"], case["time"], text_context=inp["text_context"], image_context=None)
        ref = case["out"].cuda()
        assert out.shape == ref.shape
        err = rel_err(out, ref)
        print(f"[parity] noctx_{H}x{W}: rel_err_of_range={err:.4g}")
        assert err < 1e-2, (case["shape"], err)
    ode = g["ode"]
    B, H, W = ode["shape"]
    inp = {k: v.cuda() for k, v in synthetic.restoration_inputs(B, H, W, T=ode["T"], seed=ode["seed"]).items()}
    sde = IRSDE(max_sigma=50, T=ode["T"], schedule="cosine", eps=0.005, device=cuda)
    sde.set_model(m)
    sde.set_mu(inp["lq"])
    x_T = inp["lq"] + inp["eps0"] * sde.max_sigma
    got = sde.reverse_ode(x_T)
    ref = ode["out"].cuda()
    d = (got.clamp(0, 1) - ref.clamp(0, 1)).abs().max().item()
    print(f"[parity] noctx_reverse_ode_T100: max_abs_clamped={d:.4g}, ref_absmax={ref.abs().max().item():.4g}")
    assert d <= 2e-2


@pytest.mark.parametrize("mode", ["sde", "posterior"])
def test_reduced_step_sampling_vs_reference_golden(model, cuda, mode):
    """IRSDE(T=100, sample_T=20) through the drop-in sampler (fused loop) against the reference's loop."""
    from daclip_b200 import synthetic
    from daclip_b200.sde import IRSDE
    m, _, _ = model
    g = torch.load(os.path.join(GOLD, "sampler_reduced.pt"), weights_only=False)
    ST = g["sample_T"]
    inp = {k: v.cuda() for k, v in synthetic.restoration_inputs(1, 32, 32, T=ST, seed=g["seed"]).items()}
    sde = IRSDE(max_sigma=50, T=g["T"], sample_T=ST, schedule="cosine", eps=0.005, device=cuda)
    assert sde.sample_scale == g["T"] / ST and torch.equal(sde.sigma_bars.cpu(), g["sigma_bars"])
    sde.set_model(m)
    sde.set_mu(inp["lq"])
    x_T = inp["lq"] + inp["eps0"] * sde.max_sigma
    fn = sde.reverse_sde if mode == "sde" else sde.reverse_posterior
    out = fn(x_T, noise=inp["noise"], text_context=inp["text_context"], image_context=inp["image_context"]).cpu()
    ref = g[mode]
    assert (out - ref).abs().max().item() < 2e-2, (out - ref).abs().max().item()
    assert psnr(out, ref) > 45.0


@pytest.mark.parametrize("B,H,W", [(1, 256, 256), (5, 64, 96), (16, 128, 128)])
def test_evaluation_is_bit_reproducible(model, B, H, W):
    """No atomics on the path: the LinearAttention context is stored as per-CTA partial records and merged in a fixed
    order, so replaying the step graph on the same inputs gives the same bits - also when one image spans every CTA
    (batch 1) and when CTAs hold pieces of several images."""
    from daclip_b200 import synthetic
    m, _, _ = model
    inp = {k: v.cuda() for k, v in synthetic.restoration_inputs(B, H, W, T=1, seed=9).items()}
    outs = []
    for _ in range(3):
        outs.append(m(inp["lq"] + 0.1 * inp["eps0"], inp["lq"], 61.0, text_context=inp["text_context"],
                      image_context=inp["image_context"]).clone())
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_model_from_option_yaml_vs_reference_golden(model, gold, cuda, tmp_path):
    """The reference's entry sequence (config/daclip-sde/test.py:26-28,68-85,112-127): options.parse(test.yml) ->
    dict_to_nonedict -> create_model(opt) (checkpoint from path.pretrain_model_G, keys with the DataParallel `module.`
    prefix) -> IRSDE(**opt['sde']) -> feed_data -> test(sde, opt['sde']['sampling_mode']) - against the reference's own
    T = 100 loop (tests/golden/unet_sampler.pt)."""
    import unittest.mock as um
    from daclip_b200 import options, synthetic
    from daclip_b200.model import create_model
    from daclip_b200.sde import IRSDE
    _, sd, kw = model
    ckpt = tmp_path / "universal-ir.pth"
    torch.save({"module." + k: v for k, v in sd.items()}, ckpt)
    opt = options.parse(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "options_test.yml"),
                        is_train=False, root=str(tmp_path))
    opt["path"]["pretrain_model_G"] = str(ckpt)                 # the only edit a user makes: where the weights are
    opt = options.dict_to_nonedict(opt)
    assert dict(opt["network_G"]["setting"]) == {k: kw[k] for k in opt["network_G"]["setting"]}
    m = create_model(opt)
    assert m.device.type == "cuda"
    so = opt["sde"]
    sde = IRSDE(max_sigma=so["max_sigma"], T=so["T"], schedule=so["schedule"], eps=so["eps"], device=m.device)
    sde.set_model(m.model)
    sde.noise_source = "torch"                                  # the reference's RNG stream, patched below
    g = gold["trajectory"]
    inp = {k: v.cuda() for k, v in synthetic.restoration_inputs(1, 32, 32, T=g["T"], seed=g["seed"]).items()}
    x_T = inp["lq"] + inp["eps0"] * sde.max_sigma
    m.feed_data(x_T, inp["lq"], None, text_context=inp["text_context"], image_context=inp["image_context"])
    it = iter(inp["noise"])
    with um.patch("torch.randn_like", lambda t: next(it)):
        m.test(sde, mode=so["sampling_mode"])
    x = m.get_current_visuals(need_GT=False)["Output"]
    ref = g[so["sampling_mode"]][0]
    err = (x - ref).abs().max().item()
    assert err <= 2e-2 and psnr(x, ref) >= 45.0, (err, psnr(x, ref))


def test_default_noise_source_follows_torch_seed(model, cuda):
    """Without injected noise the fused loop draws its Gaussians inside the update kernel (Philox, seeded once per call
    from torch's generator): torch.manual_seed makes a restoration reproducible, another seed gives another sample."""
    from daclip_b200 import synthetic
    from daclip_b200.sde import IRSDE
    m, _, _ = model
    inp = {k: v.cuda() for k, v in synthetic.restoration_inputs(2, 32, 32, T=1, seed=9).items()}
    sde = IRSDE(max_sigma=50, T=100, sample_T=10, schedule="cosine", eps=0.005, device=cuda)
    sde.set_model(m)
    sde.set_mu(inp["lq"])
    x_T = inp["lq"] + inp["eps0"] * sde.max_sigma
    outs = []
    for seed in (5, 5, 6):
        torch.manual_seed(seed)
        outs.append(sde.reverse_sde(x_T, text_context=inp["text_context"], image_context=inp["image_context"]))
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1]) and not torch.equal(outs[0], outs[2])
