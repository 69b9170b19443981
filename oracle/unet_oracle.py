"""Oracle: ConditionalUNet noise predictor (fp32, torch, functional).

TEST INFRASTRUCTURE - see oracle/__init__.py.  Restates the forward pass of
/root/reference/universal-image-restoration/config/daclip-sde/models/modules/
  DenoisingUNet_arch.py  (cited as arch.py:<line>)
  module_util.py         (cited as mu.py:<line>)
  attention.py           (cited as attn.py:<line>)
as pure functions over a state dict in the reference's own key/shape layout
(224 tensors for the test.yml setting).  No nn.Module, no einops: every block
is spelled with conv2d / linear / matmul so each CUDA kernel has a one-to-one
oracle function.
"""
import math

import torch
import torch.nn.functional as F


class UNetConfig:
    """Constructor arguments of ConditionalUNet (arch.py:22-23) + derived dims."""

    def __init__(self, in_nc=3, out_nc=3, nf=64, ch_mult=(1, 2, 4, 8), context_dim=512,
                 use_degra_context=True, use_image_context=True, upscale=1, scale=None):
        # `scale` is the wild-ir variant's argument (config/wild-ir/models/modules/DenoisingUNet_arch.py:22-40): 0.5 adds
        # a Downsample(nf, nf) after init_conv and an Upsample(nf, nf) before the final concat
        self.wild = scale is not None                      # only the wild-ir class takes `scale`
        self.scale = 1 if scale is None else scale
        self.in_nc, self.out_nc, self.nf = in_nc, out_nc, nf
        self.ch_mult = list(ch_mult)
        self.depth = len(self.ch_mult)
        self.context_dim = -1 if context_dim is None else context_dim
        self.use_degra_context = use_degra_context
        self.use_image_context = use_image_context
        mult = [1] + self.ch_mult
        self.dims = [(nf * mult[i], nf * mult[i + 1]) for i in range(self.depth)]
        self.mid_dim = nf * mult[-1]
        self.time_dim = nf * 4

    @property
    def spatial_transformer(self):
        return self.use_image_context and self.context_dim > 0

    def level_is_transformer(self, i):
        # arch.py:77-82: SpatialTransformer only at i >= 3, LinearAttention elsewhere; the wild-ir class tests
        # i >= depth - 1 instead (config/wild-ir/models/modules/DenoisingUNet_arch.py:83-84)
        return self.spatial_transformer and i >= (self.depth - 1 if self.wild else 3)


USE_SDPA = False     # see attention(): set by bench.py's library baseline for one of its legs


def silu(x):
    return x * torch.sigmoid(x)


def time_embedding(sd, time, nf):
    """SinusoidalPosEmb -> Linear -> GELU -> Linear (arch.py:51-56, mu.py:36-48)."""
    half = nf // 2
    freq = torch.exp(torch.arange(half, device=time.device) * -(math.log(10000) / (half - 1)))
    arg = time[:, None] * freq[None, :]
    emb = torch.cat([arg.sin(), arg.cos()], dim=-1)
    h = F.gelu(F.linear(emb, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"]))
    return F.linear(h, sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])


def prompt_embedding(sd, text_context):
    """softmax(text_mlp(ctx)) * prompt -> prompt_mlp (arch.py:134-137)."""
    h = silu(F.linear(text_context, sd["text_mlp.0.weight"], sd["text_mlp.0.bias"]))
    h = F.linear(h, sd["text_mlp.2.weight"], sd["text_mlp.2.bias"])
    h = torch.softmax(h, dim=1) * sd["prompt"]
    return F.linear(h, sd["prompt_mlp.weight"], sd["prompt_mlp.bias"])


def res_block(sd, p, x, t_emb):
    """ResBlock (mu.py:132-153): conv3x3 -> FiLM -> SiLU -> conv3x3 -> SiLU, + skip."""
    film = F.linear(silu(t_emb), sd[p + "mlp.1.weight"], sd[p + "mlp.1.bias"])
    scale, shift = film[:, :, None, None].chunk(2, dim=1)
    h = F.conv2d(x, sd[p + "block1.proj.weight"], padding=1)
    h = silu(h * (scale + 1) + shift)
    h = silu(F.conv2d(h, sd[p + "block2.proj.weight"], padding=1))
    key = p + "res_conv.weight"
    return h + (F.conv2d(x, sd[key]) if key in sd else x)


def channel_layernorm(x, g, eps=1e-5):
    """Gain-only LayerNorm over channels, biased variance (mu.py:77-86)."""
    mean = x.mean(dim=1, keepdim=True)
    var = x.var(dim=1, unbiased=False, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * g


def linear_attention(sd, p, x, heads=4, dim_head=32):
    """LinearAttention (mu.py:157-185): q softmax over head channels, k softmax over pixels."""
    b, c, h, w = x.shape
    n = h * w
    qkv = F.conv2d(x, sd[p + "to_qkv.weight"])
    q, k, v = [t.reshape(b, heads, dim_head, n) for t in qkv.chunk(3, dim=1)]
    q = torch.softmax(q, dim=2) * dim_head ** -0.5
    k = torch.softmax(k, dim=3)
    v = v / n
    ctx = torch.matmul(k, v.transpose(2, 3))            # [b, heads, d, e]
    out = torch.matmul(ctx.transpose(2, 3), q)          # [b, heads, e, n]
    out = out.reshape(b, heads * dim_head, h, w)
    out = F.conv2d(out, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])
    return channel_layernorm(out, sd[p + "to_out.1.g"])


def attention(sd, p, x, context, heads):
    """CrossAttention (attn.py:152-193); context=None means self-attention."""
    b, n, _ = x.shape
    ctx = x if context is None else context
    q = F.linear(x, sd[p + "to_q.weight"])
    k = F.linear(ctx, sd[p + "to_k.weight"])
    v = F.linear(ctx, sd[p + "to_v.weight"])
    d = q.shape[-1] // heads

    def split(t):
        return t.reshape(b, t.shape[1], heads, d).transpose(1, 2)   # [b, heads, len, d]

    q, k, v = split(q), split(k), split(v)
    if USE_SDPA:     # bench.py's strongest library leg only (fused attention of the installed PyTorch); never in parity tests
        out = F.scaled_dot_product_attention(q, k, v)
    else:
        sim = torch.matmul(q, k.transpose(2, 3)) * d ** -0.5
        out = torch.matmul(torch.softmax(sim, dim=-1), v)
    out = out.transpose(1, 2).reshape(b, n, heads * d)
    return F.linear(out, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])


def transformer_block(sd, p, x, context, heads):
    """BasicTransformerBlock (attn.py:196-215): self-attn, cross-attn, GEGLU FFN, all pre-LN residual."""
    c = x.shape[-1]

    def ln(t, name):
        return F.layer_norm(t, (c,), sd[p + name + ".weight"], sd[p + name + ".bias"], 1e-5)

    x = attention(sd, p + "attn1.", ln(x, "norm1"), None, heads) + x
    x = attention(sd, p + "attn2.", ln(x, "norm2"), context, heads) + x
    y = F.linear(ln(x, "norm3"), sd[p + "ff.net.0.proj.weight"], sd[p + "ff.net.0.proj.bias"])
    val, gate = y.chunk(2, dim=-1)                      # attn.py:43-44
    y = F.linear(val * F.gelu(gate), sd[p + "ff.net.2.weight"], sd[p + "ff.net.2.bias"])
    return y + x


def spatial_transformer(sd, p, x, context, heads):
    """SpatialTransformer (attn.py:218-261): GN32 -> 1x1 -> block -> 1x1 -> + input."""
    b, c, h, w = x.shape
    y = F.group_norm(x, 32, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    y = F.conv2d(y, sd[p + "proj_in.weight"], sd[p + "proj_in.bias"])
    inner = y.shape[1]
    y = y.reshape(b, inner, h * w).transpose(1, 2)
    y = transformer_block(sd, p + "transformer_blocks.0.", y, context, heads)
    y = y.transpose(1, 2).reshape(b, inner, h, w)
    y = F.conv2d(y, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    return y + x


def attn_layer(sd, p, x, context, is_transformer):
    """Residual(PreNorm(dim, attn)) (mu.py:27-33,89-97)."""
    y = channel_layernorm(x, sd[p + "fn.norm.g"])
    if is_transformer:
        y = spatial_transformer(sd, p + "fn.fn.", y, context, heads=x.shape[1] // 32)
    else:
        y = linear_attention(sd, p + "fn.fn.", y)
    return y + x


def unet_forward(sd, cfg: UNetConfig, xt, cond, time, text_context=None, image_context=None,
                 taps=None):
    """ConditionalUNet.forward (arch.py:118-174).  ``taps``: optional dict that
    receives named intermediate activations (used to localise kernel bugs)."""
    if isinstance(time, (int, float)):
        time = torch.tensor([time], device=xt.device)
    time = time.to(torch.float32)

    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    x = torch.cat([xt - cond, cond], dim=1)
    H, W = x.shape[2:]
    s = 2 ** cfg.depth                                     # arch.py:111-116
    ph, pw = (s - H % s) % s, (s - W % s) % s
    if ph or pw:
        x = F.pad(x, (0, pw, 0, ph), mode="reflect")
    x = tap("init_conv", F.conv2d(x, sd["init_conv.weight"], padding=3))
    x_first = x
    if cfg.scale == 0.5:                                   # wild-ir arch.py:136-140
        x = tap("downsample", F.conv2d(x, sd["downsample.weight"], sd["downsample.bias"], stride=2, padding=1))

    t = time_embedding(sd, time, cfg.nf)
    context = None
    if cfg.context_dim > 0:
        if cfg.use_degra_context and text_context is not None:
            t = t + prompt_embedding(sd, text_context)
        if cfg.use_image_context and image_context is not None:
            context = image_context[:, None, :]
    tap("t_emb", t)

    skips = []
    for i in range(cfg.depth):
        p = f"downs.{i}."
        x = tap(p + "0", res_block(sd, p + "0.", x, t))
        skips.append(x)
        x = tap(p + "1", res_block(sd, p + "1.", x, t))
        x = tap(p + "2", attn_layer(sd, p + "2.", x, context, cfg.level_is_transformer(i)))
        skips.append(x)
        if i != cfg.depth - 1:
            x = F.conv2d(x, sd[p + "3.weight"], sd[p + "3.bias"], stride=2, padding=1)
        else:
            x = F.conv2d(x, sd[p + "3.weight"], padding=1)
        tap(p + "3", x)

    x = tap("mid_block1", res_block(sd, "mid_block1.", x, t))
    x = tap("mid_attn", attn_layer(sd, "mid_attn.", x, context, cfg.spatial_transformer))
    x = tap("mid_block2", res_block(sd, "mid_block2.", x, t))

    for j in range(cfg.depth):
        i = cfg.depth - 1 - j                              # ups were insert(0, ...) (arch.py:91)
        p = f"ups.{j}."
        x = tap(p + "0", res_block(sd, p + "0.", torch.cat([x, skips.pop()], dim=1), t))
        x = tap(p + "1", res_block(sd, p + "1.", torch.cat([x, skips.pop()], dim=1), t))
        x = tap(p + "2", attn_layer(sd, p + "2.", x, context, cfg.level_is_transformer(i)))
        if i != 0:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = F.conv2d(x, sd[p + "3.1.weight"], sd[p + "3.1.bias"], padding=1)
        else:
            x = F.conv2d(x, sd[p + "3.weight"], padding=1)
        tap(p + "3", x)

    if cfg.scale == 0.5:                                   # wild-ir arch.py:176-180
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        x = tap("upsample", F.conv2d(x, sd["upsample.1.weight"], sd["upsample.1.bias"], padding=1))
    x = res_block(sd, "final_res_block.", torch.cat([x, x_first], dim=1), t)
    tap("final_res_block", x)
    x = F.conv2d(x, sd["final_conv.weight"], sd["final_conv.bias"], padding=1)
    return x[..., :H, :W].contiguous()


def make_denoiser(sd, cfg):
    """Adapter with the call shape the sampler expects: model(x, mu, t, **ctx) (sde_utils.py:197)."""
    def denoiser(x, mu, t, text_context=None, image_context=None):
        return unet_forward(sd, cfg, x, mu, t, text_context, image_context)
    return denoiser
