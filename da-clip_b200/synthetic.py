"""Seeded synthetic weights and inputs in the reference's state-dict layouts.

No pretrained checkpoint exists offline (options/test.yml:43-44 point at absent files), so parity and
throughput use random weights that every party can regenerate from a seed: the golden generator (which loads
them into the REFERENCE modules), the oracle, the tests and bench.py.  Values come from a CPU torch.Generator
in state-dict key order, so they are identical on every host with the same torch build.  All-zero-initialised
reference modules (ControlTransformer.zero_modules, SpatialTransformer.proj_out) get non-zero values on
purpose - otherwise the control path and every transformer block would be numerically dead and untested.
"""
import math

import torch


def _is_gain(name):
    parts = name.split(".")
    last, parent = parts[-1], (parts[-2] if len(parts) > 1 else "")
    return last == "g" or (last == "weight" and (parent.startswith("norm") or parent.startswith("ln_")))


def _fill(name, t, g):
    shape = t.shape
    if _is_gain(name):
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if name == "prompt":
        return torch.rand(shape, generator=g)
    if name.endswith("bias") or name.endswith("in_proj_bias"):
        return 0.05 * torch.randn(shape, generator=g)
    if name.endswith(("class_embedding", "positional_embedding")):
        return shape[-1] ** -0.5 * torch.randn(shape, generator=g)
    if name.endswith("proj") and len(shape) == 2:          # ViT output projection [width, embed]
        return shape[0] ** -0.5 * torch.randn(shape, generator=g)
    if name == "logit_scale" or len(shape) == 0:
        return torch.tensor(math.log(1 / 0.07))
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    return fan_in ** -0.5 * torch.randn(shape, generator=g)


def randomize_state_dict(shapes, seed):
    """shapes: ordered {key: shape-or-tensor}.  Returns {key: fp32 CPU tensor}."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in shapes.items():
        t = v if torch.is_tensor(v) else torch.empty(v)
        if not t.is_floating_point():
            out[k] = t.clone()
            continue
        out[k] = _fill(k, t, g).to(torch.float32).reshape(t.shape)
    return out


def unet_state_dict(seed=0, stabilize=True, **ctor):
    """Synthetic ConditionalUNet weights (test.yml setting by default) in the reference key order."""
    from .unet import ConditionalUNet
    kw = dict(in_nc=3, out_nc=3, nf=64, ch_mult=[1, 2, 4, 8], context_dim=512, use_degra_context=True,
              use_image_context=True)
    kw.update(ctor)
    m = ConditionalUNet(**kw)
    sd = randomize_state_dict({k: v.shape for k, v in m.state_dict().items()}, seed)
    if stabilize:
        _add_denoiser_path(sd, kw["nf"])
    if "prompt" in sd:
        _amplify_prompt_path(sd)
    return sd, kw


def _amplify_prompt_path(sd, prompt_gain=64.0, logit_gain=4.0):
    """Make the degradation-prompt path numerically visible.  With default-style random weights
    `softmax(text_mlp(text_context)) * prompt` (arch.py:134-137) is ~1/256 per entry, so swapping text_context moved a
    64x64 prediction by 6e-4 (image_context: 0.19) and a wiring bug on that path would pass every end-to-end check.
    A peakier softmax (x4 logits) and a larger learned prompt (x64) give the path the weight a trained prompt has:
    swapping text_context now moves the prediction by ~0.4 (measured with the oracle; tests assert >= 5e-2)."""
    sd["prompt"] *= prompt_gain
    sd["text_mlp.2.weight"] *= logit_gain
    sd["text_mlp.2.bias"] *= logit_gain


def _add_denoiser_path(sd, nf, gain=5.1, random_scale=0.05):
    """Make the random network behave like a (crude) denoiser so that T-step trajectories stay O(1).

    With an untrained net the reverse process EXPANDS x - mu by 1/eps = 200x over T steps (it inverts the
    forward contraction), which turns any end-to-end tolerance into noise.  The ideal prediction for x0 = mu is
    n = (x - mu) / sigma_bar_t; sigma_bar_t ~ 0.196 for most steps, so a fixed linear path
    n_lin = gain * (xt - cond), gain = 1/0.196, cancels the expansion.  The architecture has such a path:
    init_conv -> (concat x_) -> final_res_block.res_conv -> final_conv.  The randomly initialised remainder of
    the network still contributes ~0.15 std on top, through every layer."""
    w = sd["init_conv.weight"]
    w[0:3] = 0
    for c in range(3):
        w[c, c, 3, 3] = 1.0                               # x_[c] = (xt - cond)[c]
    r = sd["final_res_block.res_conv.weight"]
    r[0:3] = 0
    for c in range(3):
        r[c, nf + c, 0, 0] = 1.0                          # skip[c] = x_[c]
    sd["final_res_block.block2.proj.weight"][0:3] = 0     # SiLU(0) = 0: nothing else lands on channels 0..2
    sd["final_conv.weight"] *= random_scale
    sd["final_conv.bias"] *= random_scale
    f = sd["final_conv.weight"]
    f[:, 0:3] = 0
    for c in range(3):
        f[c, c, 1, 1] = gain


def vit_tower_shapes(prefix, width=768, layers=12, patch=32, grid=7, embed=512, control=False):
    s = {}
    s[prefix + "class_embedding"] = (width,)
    s[prefix + "positional_embedding"] = (grid * grid + 1, width)
    s[prefix + "proj"] = (width, embed)
    s[prefix + "conv1.weight"] = (width, 3, patch, patch)
    s[prefix + "ln_pre.weight"] = (width,)
    s[prefix + "ln_pre.bias"] = (width,)
    tp = prefix + ("transformer.transformer." if control else "transformer.")
    for i in range(layers):
        b = f"{tp}resblocks.{i}."
        s[b + "ln_1.weight"] = (width,); s[b + "ln_1.bias"] = (width,)
        s[b + "attn.in_proj_weight"] = (3 * width, width); s[b + "attn.in_proj_bias"] = (3 * width,)
        s[b + "attn.out_proj.weight"] = (width, width); s[b + "attn.out_proj.bias"] = (width,)
        s[b + "ln_2.weight"] = (width,); s[b + "ln_2.bias"] = (width,)
        s[b + "mlp.c_fc.weight"] = (4 * width, width); s[b + "mlp.c_fc.bias"] = (4 * width,)
        s[b + "mlp.c_proj.weight"] = (width, 4 * width); s[b + "mlp.c_proj.bias"] = (width,)
    if control:
        for i in range(layers):
            s[f"{prefix}transformer.zero_modules.{i}.weight"] = (width, width)
            s[f"{prefix}transformer.zero_modules.{i}.bias"] = (width,)
    s[prefix + "ln_post.weight"] = (width,)
    s[prefix + "ln_post.bias"] = (width,)
    return s


VIT_ARCHS = {"ViT-B-32": dict(width=768, layers=12, patch=32, grid=7, embed=512),
             "ViT-L-14": dict(width=1024, layers=24, patch=14, grid=16, embed=768)}


def daclip_visual_state_dict(seed=10, arch="ViT-B-32"):
    """Synthetic weights of the two ViT towers used by encode_image(control=True):
    `visual.*` (frozen CLIP tower; the reference aliases it as `clip.visual.*`) and `visual_control.*`."""
    shapes = {}
    shapes.update(vit_tower_shapes("visual.", **VIT_ARCHS[arch]))
    shapes.update(vit_tower_shapes("visual_control.", control=True, **VIT_ARCHS[arch]))
    sd = randomize_state_dict(shapes, seed)
    # keep the control signal a perturbation, as a trained ControlNet-style branch would be
    for k in sd:
        if "zero_modules" in k:
            sd[k] = sd[k] * 0.25
    return sd


TEXT_ARCHS = {"ViT-B-32": dict(width=512, layers=12, embed=512), "ViT-L-14": dict(width=768, layers=12, embed=768)}


def daclip_text_state_dict(seed=12, arch="ViT-B-32", context_length=77, vocab_size=49408):
    """Synthetic weights of the CLIP text tower under the reference's `clip.*` keys (open_clip/model.py:203-213)."""
    a = TEXT_ARCHS[arch]
    w = a["width"]
    s = {"clip.positional_embedding": (context_length, w), "clip.text_projection": (w, a["embed"])}
    for i in range(a["layers"]):
        b = f"clip.transformer.resblocks.{i}."
        s[b + "ln_1.weight"] = (w,); s[b + "ln_1.bias"] = (w,)
        s[b + "attn.in_proj_weight"] = (3 * w, w); s[b + "attn.in_proj_bias"] = (3 * w,)
        s[b + "attn.out_proj.weight"] = (w, w); s[b + "attn.out_proj.bias"] = (w,)
        s[b + "ln_2.weight"] = (w,); s[b + "ln_2.bias"] = (w,)
        s[b + "mlp.c_fc.weight"] = (4 * w, w); s[b + "mlp.c_fc.bias"] = (4 * w,)
        s[b + "mlp.c_proj.weight"] = (w, 4 * w); s[b + "mlp.c_proj.bias"] = (w,)
    s["clip.token_embedding.weight"] = (vocab_size, w)
    s["clip.ln_final.weight"] = (w,)
    s["clip.ln_final.bias"] = (w,)
    sd = randomize_state_dict(s, seed)
    sd["clip.token_embedding.weight"] *= 8.0       # O(1) token rows, so the positional term does not dominate
    return sd


def restoration_inputs(B, H, W, T=100, seed=1, ctx_dim=512):
    """LQ image batch in [0,1], the noisy start state, per-step Gaussian draws and DA-CLIP-like contexts."""
    g = torch.Generator().manual_seed(seed)
    lq = torch.rand(B, 3, H, W, generator=g)
    eps0 = torch.randn(B, 3, H, W, generator=g)
    noise = torch.randn(T, B, 3, H, W, generator=g)
    text_ctx = torch.randn(B, ctx_dim, generator=g)
    image_ctx = torch.randn(B, ctx_dim, generator=g)
    return dict(lq=lq, eps0=eps0, noise=noise, text_context=text_ctx, image_context=image_ctx)


def natural_image(h, w, seed=0):
    """A seeded float32 HWC RGB test image in [0, 1] with both smooth structure and pixel-level detail (so that a
    resampler's antialiasing taps and rounding all matter), as a numpy array - the form the reference's dataset hands
    to clip_transform (data/LQGT_dataset.py:139-143)."""
    import numpy as np
    g = torch.Generator().manual_seed(seed)
    yy = torch.linspace(0, 1, h).view(h, 1, 1)
    xx = torch.linspace(0, 1, w).view(1, w, 1)
    ph = torch.rand(3, generator=g).view(1, 1, 3) * 6.28
    base = 0.5 + 0.35 * torch.sin(9.0 * xx + 5.0 * yy + ph) * torch.cos(7.0 * yy - 3.0 * xx + ph)
    img = (base + 0.15 * torch.randn(h, w, 3, generator=g)).clamp(0, 1)
    return np.ascontiguousarray(img.numpy().astype(np.float32))
