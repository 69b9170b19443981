"""GPU parity at the BENCHMARKED configurations (BASELINE.json configs C2 / C4 / C5) and of the pieces round 1 left
unpinned: full T = 100 restorations through `DenoisingModel.feed_data / test` at batch 16 256^2 (both samplers) and
batch 8 512^2 against the fp32 oracle on the same device (TF32 off), `IRSDE.noise_state` against the reference
expression under the same seeded RNG, the batched driver against the ORACLE chain (not against itself), the
reference's 128x128 trajectory golden, the prompt path's sensitivity, and a strict degradation-argmax count.

Gate (north_star): per restored image max-abs <= 2e-2 on the clamped [0, 1] image and PSNR >= 45 dB; argmax bit-exact.
Per-step Gaussian noise is injected the way the goldens were made from the reference: torch.randn_like is patched to
hand out identical pre-generated tensors on both sides.  Achieved numbers are printed (pytest -s) and written to
gpurun_out/parity_configs.json when that directory exists; DESIGN.md quotes them.
"""
import json
import math
import os
import unittest.mock as um

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
_REPORT = {}


def _record(key, **vals):
    _REPORT[key] = vals
    print(f"[parity] {key}: " + ", ".join(f"{k}={v:.4g}" if isinstance(v, float) else f"{k}={v}" for k, v in vals.items()))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_configs.json"), "w") as f:
            json.dump(_REPORT, f, indent=1, sort_keys=True)


def per_image_gate(x, ref):
    """(max-abs, min PSNR) over the batch on the clamped [0, 1] images (what tensor2img quantises)."""
    a, b = x.clamp(0, 1).float(), ref.clamp(0, 1).float()
    worst_abs = (a - b).abs().flatten(1).max(dim=1).values
    mse = ((a - b) ** 2).flatten(1).mean(dim=1)
    psnr = torch.where(mse > 0, 10 * torch.log10(1.0 / mse.clamp_min(1e-30)), torch.full_like(mse, 99.0))
    return worst_abs.max().item(), psnr.min().item(), (x - ref).abs().max().item()


@pytest.fixture(scope="module")
def stack(cuda):
    from daclip_b200 import synthetic
    from daclip_b200.model import create_model
    from daclip_b200.sde import IRSDE
    from oracle import sde_oracle as S
    from oracle import unet_oracle as O
    sd, kw = synthetic.unet_state_dict(0)
    opt = {"gpu_ids": [0], "is_train": False, "dist": False, "model": "denoising",
           "network_G": {"which_model_G": "ConditionalUNet", "setting": dict(kw)},
           "path": {"pretrain_model_G": None, "strict_load": True}}
    model = create_model(opt)
    model.load_state_dict_into_model(sd)
    sde = IRSDE(max_sigma=50, T=100, schedule="cosine", eps=0.005, device=cuda)
    sde.noise_source = "torch"          # the reference's RNG stream (T x torch.randn_like): what the tests below patch
    sde.set_model(model.model)
    sdc = {k: v.to(cuda) for k, v in sd.items()}
    den = O.make_denoiser(sdc, O.UNetConfig(**kw))
    sch = S.Schedule(50, 100, "cosine", 0.005)
    return model, sde, den, sch


def _device_inputs(B, H, W, T, seed, cuda):
    """Seeded inputs generated on the device (a [100, 8, 3, 512, 512] noise tensor is 2.5 GB: too slow from the host)."""
    g = torch.Generator(device=cuda).manual_seed(seed)
    lq = torch.rand(B, 3, H, W, device=cuda, generator=g)
    eps0 = torch.randn(B, 3, H, W, device=cuda, generator=g)
    noise = torch.randn(T, B, 3, H, W, device=cuda, generator=g)
    text = torch.randn(B, 512, device=cuda, generator=g)
    image = torch.randn(B, 512, device=cuda, generator=g)
    return lq, eps0, noise, text, image


@pytest.mark.parametrize("name,B,S,mode", [("C2", 16, 256, "posterior"), ("C2", 16, 256, "sde"),
                                           ("C4", 8, 512, "posterior")])
def test_full_restoration_at_benchmark_config(stack, cuda, name, B, S, mode):
    """feed_data -> test(sde, mode) -> output for the whole benchmarked batch, all T = 100 steps, vs the oracle loop
    (sde_utils.py:261-313 driving arch.py:118-174) in fp32 on the GPU with identical noise."""
    from oracle import sde_oracle as So
    model, sde, den, sch = stack
    T = 100
    lq, eps0, noise, text, image = _device_inputs(B, S, S, T, seed=1000 + S + (mode == "sde"), cuda=cuda)
    x_T = lq + eps0 * sde.max_sigma
    model.feed_data(x_T, lq, None, text_context=text, image_context=image)
    it = iter(noise)
    with um.patch("torch.randn_like", lambda t: next(it)):
        model.test(sde, mode=mode)
    got = model.output
    assert next(it, None) is None, "the loop must consume exactly T noise tensors"
    with torch.no_grad():
        ref = So.reverse(sch, den, x_T, lq, mode=mode, noise=noise, text_context=text, image_context=image)
    worst, psnr, raw = per_image_gate(got, ref)
    _record(f"{name}_{S}_b{B}_{mode}_T{T}", max_abs_clamped=worst, min_psnr_db=psnr, max_abs_unclamped=raw,
            ref_absmax=ref.abs().max().item())
    assert torch.isfinite(got).all()
    assert worst <= 2e-2, f"max-abs {worst:.4g} over the batch"
    assert psnr >= 45.0, f"min PSNR {psnr:.2f} dB over the batch"


@pytest.mark.parametrize("mode", ["sde", "posterior"])
def test_trajectory_128_vs_reference_golden(stack, cuda, mode):
    """The REFERENCE's own T = 100 loops at 128x128 (tests/golden/trajectory_128.pt, made by oracle/gen_golden.py):
    every level runs its production kernel (tcgen05 k|v and q-out kernels, 256-token tensor-core attention)."""
    from daclip_b200 import synthetic
    model, sde, _, _ = stack
    g = torch.load(os.path.join(GOLD, "trajectory_128.pt"), weights_only=False)
    B, H, W = g["shape"]
    inp = {k: v.to(cuda) for k, v in synthetic.restoration_inputs(B, H, W, T=g["T"], seed=g["seed"]).items()}
    x_T = inp["lq"] + inp["eps0"] * sde.max_sigma
    model.feed_data(x_T, inp["lq"], None, text_context=inp["text_context"], image_context=inp["image_context"])
    it = iter(inp["noise"])
    with um.patch("torch.randn_like", lambda t: next(it)):
        model.test(sde, mode=mode)
    worst, psnr, raw = per_image_gate(model.output.cpu(), g[mode])
    _record(f"ref128_{mode}", max_abs_clamped=worst, min_psnr_db=psnr, max_abs_unclamped=raw)
    assert worst <= 2e-2 and psnr >= 45.0, (worst, psnr)


def test_single_evaluation_error_budget(stack, cuda):
    """One denoiser evaluation at the benchmarked shape (batch 16, 256^2): error relative to the output range, printed and
    bounded at 2x what the bf16 path achieves (round 1 allowed 2e-2 of the range without knowing the achieved value)."""
    model, sde, den, _ = stack
    lq, eps0, _, text, image = _device_inputs(16, 256, 256, 1, seed=77, cuda=cuda)
    net = model.model.module
    worst = 0.0
    for t in (100.0, 57.0, 1.0):
        xt = lq + eps0 * (50 / 255)
        got = net(xt, lq, t, text_context=text, image_context=image)
        with torch.no_grad():
            ref = den(xt, lq, t, text_context=text, image_context=image)
        e = (got - ref).abs().max().item() / ref.abs().max().item()
        worst = max(worst, e)
    _record("single_eval_b16_256", rel_err_of_range=worst)
    assert worst < 1e-2, worst


def test_prompt_path_is_numerically_visible(stack, cuda):
    """Swapping ONLY text_context (the DA-CLIP degradation embedding -> prompt -> every FiLM, arch.py:134-137) must move
    the prediction by >= 5e-2, and the kernels must follow the oracle through that change - a wiring bug in the
    text / prompt path cannot hide behind the tolerance."""
    model, sde, den, _ = stack
    lq, eps0, _, text, image = _device_inputs(2, 64, 64, 1, seed=78, cuda=cuda)
    text2 = torch.randn(2, 512, device=cuda, generator=torch.Generator(device=cuda).manual_seed(5))
    xt = lq + eps0 * (50 / 255)
    net = model.model.module
    got = [net(xt, lq, 50.0, text_context=t, image_context=image) for t in (text, text2)]
    with torch.no_grad():
        ref = [den(xt, lq, 50.0, text_context=t, image_context=image) for t in (text, text2)]
    moved_ref = (ref[1] - ref[0]).abs().max().item()
    moved_got = (got[1] - got[0]).abs().max().item()
    err_of_delta = ((got[1] - got[0]) - (ref[1] - ref[0])).abs().max().item()
    _record("prompt_path", moved_ref=moved_ref, moved_got=moved_got, err_of_delta=err_of_delta)
    assert moved_ref >= 5e-2, moved_ref
    assert err_of_delta <= 0.15 * moved_ref, (err_of_delta, moved_ref)
    # without a text context the prompt branch is skipped on both sides (arch.py:134: `if ... text_context is not None`)
    g0 = net(xt, lq, 50.0, text_context=None, image_context=image)
    with torch.no_grad():
        r0 = den(xt, lq, 50.0, text_context=None, image_context=image)
    assert (g0 - r0).abs().max().item() / r0.abs().max().item() < 1e-2
    assert (r0 - ref[0]).abs().max().item() >= 5e-2


def test_noise_state_matches_reference_expression(cuda):
    """IRSDE.noise_state (sde_utils.py:374-375): `tensor + torch.randn_like(tensor) * max_sigma` under the same seeded RNG,
    bit for bit, for device and host inputs (the reference call site hands it the host LQ, test.py:119)."""
    from daclip_b200.sde import IRSDE
    sde = IRSDE(max_sigma=50, T=100, schedule="cosine", eps=0.005, device=cuda)
    for dev in (cuda, torch.device("cpu")):
        lq = torch.rand(3, 3, 37, 53, generator=torch.Generator().manual_seed(4)).to(dev)
        torch.manual_seed(123)
        got = sde.noise_state(lq)
        torch.manual_seed(123)
        ref = lq + torch.randn_like(lq) * sde.max_sigma
        assert got.device == lq.device and got.dtype == torch.float32
        assert torch.equal(got, ref), (got - ref).abs().max().item()
    # the draw really is N(0, max_sigma^2)
    big = torch.zeros(1, 3, 512, 512, device=cuda)
    z = sde.noise_state(big) / sde.max_sigma
    assert abs(z.mean().item()) < 5e-3 and abs(z.std().item() - 1.0) < 5e-3


def test_sde_step_rejects_mismatched_operands(cuda):
    from daclip_b200 import lib, ops
    x = torch.zeros(2, 3, 8, 8, device=cuda)
    with pytest.raises(lib.DacError):
        ops.sde_step(1, x, x, x, torch.zeros(1, 3, 8, 8, device=cuda), x.clone(), [0.0] * 5)     # eps too small
    with pytest.raises(lib.DacError):
        ops.sde_step(1, x, x, x.double(), x, x.clone(), [0.0] * 5)
    with pytest.raises(lib.DacError):
        ops.sde_step(0, x, x, x, None, x.clone(), [0.0] * 6)


def test_batched_driver_vs_oracle_chain(cuda):
    """SURVEY 8f N2 against the ORACLE flow of test.py:112-129 per image (numpy Pillow-exact clip_transform -> oracle
    DaCLIP -> noise_state with the injected draw -> oracle T-step loop -> oracle tensor2img), not against the product's
    own one-image path.  Also covers a size group split across chunks with per-chunk noise rows."""
    from daclip_b200 import imageio, synthetic
    from daclip_b200.daclip import DaCLIP
    from daclip_b200.driver import BatchedRestorer
    from daclip_b200.model import create_model
    from daclip_b200.sde import IRSDE
    from oracle import daclip_oracle as D
    from oracle import imageio_oracle as IO
    from oracle import sde_oracle as So
    from oracle import unet_oracle as O
    sd, kw = synthetic.unet_state_dict(0)
    opt = {"gpu_ids": [0], "is_train": False, "dist": False, "model": "denoising",
           "network_G": {"which_model_G": "ConditionalUNet", "setting": dict(kw)},
           "path": {"pretrain_model_G": None, "strict_load": True}}
    model = create_model(opt)
    model.load_state_dict_into_model(sd)
    T = 100
    sde = IRSDE(max_sigma=50, T=T, schedule="cosine", eps=0.005, device=cuda)
    sde.set_model(model.model)
    vis = synthetic.daclip_visual_state_dict(10)
    clip = DaCLIP().load_reference_state_dict(vis).to(cuda).eval()
    sizes = [(48, 64), (32, 32), (48, 64), (48, 64), (32, 32)]
    imgs = [synthetic.natural_image(h, w, seed=170 + i) for i, (h, w) in enumerate(sizes)]
    g = torch.Generator().manual_seed(15)
    draws = [torch.randn(T + 1, 1, 3, h, w, generator=g) for h, w in sizes]        # [0] = the noise_state draw

    # oracle chain, one image at a time
    sdc = {k: v.to(cuda) for k, v in sd.items()}
    visc = {k: v.to(cuda) for k, v in vis.items()}
    den = O.make_denoiser(sdc, O.UNetConfig(**kw))
    sch = So.Schedule(50, T, "cosine", 0.005)
    prototypes = torch.randn(10, 3, 224, 224, generator=g).to(cuda)
    with torch.no_grad():
        _, text = D.encode_image_control(visc, prototypes)     # ten fixed class directions in the degradation feature space
    want = []
    for im, nz in zip(imgs, draws):
        _, clip_in = IO.clip_transform(im, imageio.pil_bicubic_coeffs, imageio.resized_size)
        lq = torch.from_numpy(im).permute(2, 0, 1)[None].to(cuda)
        with torch.no_grad():
            ic, dc = D.encode_image_control(visc, torch.from_numpy(clip_in)[None].to(cuda))
            x_T = lq + nz[0].to(cuda) * sch.max_sigma
            out = So.reverse(sch, den, x_T, lq, mode="posterior", noise=nz[1:].to(cuda), text_context=dc,
                             image_context=ic)
        logits = D.degradation_logits(dc, text)
        want.append((out[0], IO.tensor2img(out[0].cpu().numpy()), int(logits.argmax(-1)[0]), logits))

    class FixedNoiseSDE:
        """Delegates to the IRSDE but feeds noise_state the injected draw of the images it is called with."""
        def __init__(self, inner, draws_by_size):
            self.inner, self.draws, self.cursor = inner, draws_by_size, {}
        def __getattr__(self, k):
            return getattr(self.inner, k)
        def noise_state(self, t):
            key = (t.shape[2], t.shape[3])
            s = self.cursor.get(key, 0)
            self.cursor[key] = s + t.shape[0]
            return self.inner._noise_state(t, self.draws[key][s:s + t.shape[0]])

    noise, first = {}, {}
    for hw in dict.fromkeys(sizes):
        sel = [nz for s, nz in zip(sizes, draws) if s == hw]
        first[hw] = torch.cat([nz[0] for nz in sel])
        noise[hw] = torch.cat([nz[1:] for nz in sel], dim=1).to(cuda)
    # max_batch = 2 splits the three 48x64 images into chunks of 2 + 1: each chunk must get ITS rows of the noise
    r = BatchedRestorer(model, FixedNoiseSDE(sde, first), clip, "posterior", max_batch=2, text_features=text)
    res = r.restore(imgs, noise=noise)
    worst_abs, worst_psnr, worst_u8 = 0.0, 99.0, 0
    for i, (o, u8, cls, logits) in enumerate(want):
        got = res[i]
        a, p, _ = per_image_gate(got["Output_tensor"][None], o[None])
        worst_abs, worst_psnr = max(worst_abs, a), min(worst_psnr, p)
        assert got["Output"].shape == u8.shape and got["Output"].dtype == np.uint8
        worst_u8 = max(worst_u8, int(np.abs(got["Output"].astype(int) - u8.astype(int)).max()))
        top2 = logits.topk(2, dim=-1).values[0]
        if (top2[0] - top2[1]).item() > 0.5:          # 100 * cosine: 5e-3 in cosine, far above the bf16 feature error
            assert got["degradation"] == cls, (i, got["degradation"], cls, logits)
    _record("driver_vs_oracle", max_abs_clamped=worst_abs, min_psnr_db=worst_psnr, max_u8_levels=worst_u8)
    assert worst_abs <= 2e-2 and worst_psnr >= 45.0
    assert worst_u8 <= math.ceil(2e-2 * 255) + 1
    with pytest.raises(ValueError):
        r.restore(imgs, noise={k: v[:, :1] for k, v in noise.items()})


@pytest.mark.parametrize("B", [16, 256])
def test_degradation_argmax_strict_count(cuda, B):
    """North_star: the degradation-type argmax is bit-exact.  A 'mixed-degradation' batch the way a trained DA-CLIP sees
    it: ten prototype images define the ten class directions (their oracle degradation features play the text features),
    every batch image is a prototype plus 30 % fresh noise.  ALL rows are compared (round 1 skipped undecided rows):
    mismatches == 0, and the count and the smallest top-2 gap are printed.  A second, unstructured batch (random images
    against the reference text tower's features, gaps down to 1e-3 of a logit) reports its count as well; there every
    disagreement must be a numerical tie (reference gap below the measured logit error), and there may be at most 2 %."""
    from daclip_b200 import synthetic
    from daclip_b200.daclip import DaCLIP
    from oracle import daclip_oracle as D
    gold = torch.load(os.path.join(GOLD, "daclip.pt"), weights_only=False)
    sd = synthetic.daclip_visual_state_dict(gold["weights_seed"])
    m = DaCLIP().load_reference_state_dict(sd).to(cuda).eval()
    sdc = {k: v.to(cuda) for k, v in sd.items()}
    g = torch.Generator(device=cuda).manual_seed(31 + B)
    protos = torch.randn(10, 3, 224, 224, device=cuda, generator=g)
    cls = torch.arange(B, device=cuda) % 10
    images = protos[cls] + 0.3 * torch.randn(B, 3, 224, 224, device=cuda, generator=g)

    def oracle_degra(x):
        with torch.no_grad():
            return torch.cat([D.encode_image_control(sdc, x[i:i + 32])[1] for i in range(0, x.shape[0], 32)])

    text = oracle_degra(protos)
    ref_logits = D.degradation_logits(oracle_degra(images), text)
    _, deg = m.encode_image(images, control=True)
    am, logits = m.degradation_argmax(deg, text, return_logits=True)
    ref_am = ref_logits.argmax(-1)
    top2 = ref_logits.topk(2, dim=-1).values
    mism = int((am != ref_am).sum())
    _record(f"argmax_mixed_b{B}", mismatches=mism, rows=B, min_top2_gap=(top2[:, 0] - top2[:, 1]).min().item(),
            max_logit_err=(logits - ref_logits).abs().max().item(), classes_hit=int(ref_am.unique().numel()))
    assert ref_am.unique().numel() == 10 and torch.equal(ref_am, cls)
    assert mism == 0, f"{mism} of {B} rows differ"

    # unstructured batch: random images vs the reference text tower's (random-weight) class features
    images = torch.randn(B, 3, 224, 224, device=cuda, generator=g)
    text = gold["text_features"].to(cuda)
    ref_logits = D.degradation_logits(oracle_degra(images), text)
    _, deg = m.encode_image(images, control=True)
    am, logits = m.degradation_argmax(deg, text, return_logits=True)
    ref_am = ref_logits.argmax(-1)
    err = (logits - ref_logits).abs().max().item()
    bad = (am != ref_am).nonzero().flatten().tolist()
    gaps = [(ref_logits[i, ref_am[i]] - ref_logits[i, am[i]]).item() for i in bad]
    top2 = ref_logits.topk(2, dim=-1).values
    _record(f"argmax_random_b{B}", mismatches=len(bad), rows=B, max_logit_err=err,
            min_top2_gap=(top2[:, 0] - top2[:, 1]).min().item(), worst_mismatch_gap=max(gaps, default=0.0))
    assert all(gp <= 2 * err for gp in gaps), (gaps, err)      # only numerical ties may flip
    assert len(bad) <= max(1, B // 50), (len(bad), B)
    assert err < 0.1, err                                       # 100 * cosine: 1e-3 in cosine


def test_encode_image_without_control(cuda):
    """encode_image(control=False) = the frozen CLIP tower alone (daclip_model.py:53-54) vs the reference golden."""
    from daclip_b200 import synthetic
    from daclip_b200.daclip import DaCLIP
    from oracle import daclip_oracle as D
    gold = torch.load(os.path.join(GOLD, "daclip.pt"), weights_only=False)
    sd = synthetic.daclip_visual_state_dict(gold["weights_seed"])
    m = DaCLIP().load_reference_state_dict(sd).to(cuda).eval()
    image = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(gold["image_seed"])).to(cuda)
    f = m.encode_image(image)                        # control defaults to False, as in the reference
    ref = gold["plain_features"].to(cuda)
    assert f.shape == (4, 512) and f.dtype == torch.float32
    e = (f - ref).abs().max().item() / ref.abs().max().item()
    assert e < 1e-2, e
    with torch.no_grad():
        o = D.encode_image_plain({k: v.to(cuda) for k, v in sd.items()}, image)
    assert (o - ref).abs().max().item() < 1e-3
    fn = m.encode_image(image, control=False, normalize=True)
    assert (fn.norm(dim=-1) - 1).abs().max().item() < 1e-4
    # the control=True engine of the same batch is a different plan and must still be right afterwards
    img_f, _ = m.encode_image(image, control=True)
    e2 = (img_f.cpu() - gold["image_features"]).abs().max().item() / gold["image_features"].abs().max().item()
    assert e2 < 1e-2, e2


def test_engine_cache_is_bounded(cuda):
    """one UNetEngine per (B, H, W) forever grew GPU memory with every new image size."""
    from daclip_b200 import synthetic
    from daclip_b200.unet import ConditionalUNet
    sd, kw = synthetic.unet_state_dict(0)
    m = ConditionalUNet(**kw)
    m.load_state_dict(sd)
    m = m.to(cuda).eval()
    m.MAX_ENGINES = 2
    outs = {}
    for (h, w) in [(32, 32), (32, 48), (48, 32), (32, 32)]:
        inp = {k: v.to(cuda) for k, v in synthetic.restoration_inputs(1, h, w, T=1, seed=3).items()}
        outs.setdefault((h, w), []).append(m(inp["lq"], inp["lq"], 9.0, text_context=inp["text_context"],
                                             image_context=inp["image_context"]))
        assert len(m._engines) <= 2
    assert torch.equal(outs[(32, 32)][0], outs[(32, 32)][1])          # evicted and rebuilt: same bits
    with pytest.raises(NotImplementedError):          # per-image time steps are refused, not silently truncated
        m.engine(1, 32, 32).set_time(torch.tensor([1.0, 2.0]))
    m.engine(1, 32, 32).set_time(torch.tensor([3.0, 3.0]))


def test_wrong_current_device_is_refused_or_switched(cuda):
    """launches target the CURRENT device; entry points switch to their tensors' device, raw ops refuse."""
    from daclip_b200 import lib, ops
    if torch.cuda.device_count() < 2:
        x = torch.zeros(8, device=cuda)
        lib.require_cuda(x)                    # single GPU: nothing to mix up
        return
    x = torch.zeros(1, 3, 8, 8, device="cuda:1")
    with torch.cuda.device(0):
        with pytest.raises(lib.DacError):
            ops.noise_state(x, x, x.clone(), 0.1)
