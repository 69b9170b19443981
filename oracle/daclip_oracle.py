"""Oracle: DaCLIP.encode_image(control=True) + degradation-type argmax (fp32, torch).

TEST INFRASTRUCTURE - see oracle/__init__.py.  Restates, over the reference's
631-key DaCLIP state dict,
  /root/reference/universal-image-restoration/open_clip/daclip_model.py:46-53
  /root/reference/universal-image-restoration/open_clip/transformer.py
      VisionTransformer.forward :507-555, ControlTransformer.forward :308-325,
      Transformer.forward :355-369, ResidualAttentionBlock :189-244 (cited as tr.py)
  /root/reference/da-clip/src/evaluate_daclip.py:42-50,77-84 (argmax recipe)
"""
import torch
import torch.nn.functional as F


class ViTConfig:
    """vision_cfg of model_configs/daclip_ViT-B-32.json."""

    def __init__(self, image_size=224, patch=32, width=768, layers=12, heads=12, embed_dim=512):
        self.image_size, self.patch, self.width = image_size, patch, width
        self.layers, self.heads, self.embed_dim = layers, heads, embed_dim
        self.grid = image_size // patch
        self.tokens = self.grid * self.grid + 1


def _stem(sd, p, cfg, image):
    """Patch conv, class token, positional embedding, ln_pre (tr.py:518-530)."""
    b = image.shape[0]
    x = F.conv2d(image, sd[p + "conv1.weight"], stride=cfg.patch)          # [b, width, g, g]
    x = x.reshape(b, cfg.width, -1).transpose(1, 2)                         # [b, g*g, width]
    cls = sd[p + "class_embedding"].expand(b, 1, cfg.width)
    x = torch.cat([cls, x], dim=1) + sd[p + "positional_embedding"]
    return F.layer_norm(x, (cfg.width,), sd[p + "ln_pre.weight"], sd[p + "ln_pre.bias"], 1e-5)


def _res_attn_block(sd, p, cfg, x):
    """ResidualAttentionBlock (tr.py:232-244): pre-LN MHA and pre-LN erf-GELU MLP, tokens-major [b, L, w]."""
    b, L, w = x.shape
    hd = w // cfg.heads
    y = F.layer_norm(x, (w,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
    qkv = F.linear(y, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
    q, k, v = [t.reshape(b, L, cfg.heads, hd).transpose(1, 2) for t in qkv.chunk(3, dim=-1)]
    att = torch.softmax(torch.matmul(q, k.transpose(2, 3)) * hd ** -0.5, dim=-1)
    y = torch.matmul(att, v).transpose(1, 2).reshape(b, L, w)
    x = x + F.linear(y, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
    y = F.layer_norm(x, (w,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
    y = F.gelu(F.linear(y, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]))
    return x + F.linear(y, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])


def _pool(sd, p, cfg, x):
    """Class-token pool, ln_post, projection (tr.py:543-547)."""
    pooled = F.layer_norm(x[:, 0], (cfg.width,), sd[p + "ln_post.weight"], sd[p + "ln_post.bias"], 1e-5)
    return pooled @ sd[p + "proj"]


def encode_image_control(sd, image, cfg: ViTConfig = None, taps=None):
    """(image_features, degra_features), both un-normalised [B, embed_dim] (daclip_model.py:46-53)."""
    cfg = cfg or ViTConfig()
    # control tower: every block output goes through its zero-init linear (tr.py:312-321)
    pc = "visual_control."
    x = _stem(sd, pc, cfg, image)
    hiddens = []
    for i in range(cfg.layers):
        x = _res_attn_block(sd, f"{pc}transformer.transformer.resblocks.{i}.", cfg, x)
        hiddens.append(F.linear(x, sd[f"{pc}transformer.zero_modules.{i}.weight"],
                                sd[f"{pc}transformer.zero_modules.{i}.bias"]))
    degra = _pool(sd, pc, cfg, x)
    if taps is not None:
        taps["hiddens"] = hiddens
    # frozen CLIP tower: `x += control.pop()` after block i consumes the list from
    # its END, i.e. block i receives hidden[layers-1-i] (tr.py:367-368).
    pv = "visual."
    x = _stem(sd, pv, cfg, image)
    for i in range(cfg.layers):
        x = _res_attn_block(sd, f"{pv}transformer.resblocks.{i}.", cfg, x)
        x = x + hiddens[cfg.layers - 1 - i]
    return _pool(sd, pv, cfg, x), degra


def encode_image_plain(sd, image, cfg: ViTConfig = None):
    """encode_image(control=False) = CLIP.encode_image: the frozen tower alone (daclip_model.py:53-54, model.py:232-235)."""
    cfg = cfg or ViTConfig()
    pv = "visual."
    x = _stem(sd, pv, cfg, image)
    for i in range(cfg.layers):
        x = _res_attn_block(sd, f"{pv}transformer.resblocks.{i}.", cfg, x)
    return _pool(sd, pv, cfg, x)


def encode_text(sd, text, heads=8, prefix="clip."):
    """CLIP.encode_text (open_clip/model.py:237-249): token + positional embedding, the ResidualAttentionBlock chain
    under the causal mask of build_attention_mask (tr.py:629-635), ln_final, the row of each prompt's end-of-text token
    (= argmax of the ids) times text_projection.  text: int64 [N, 77]; un-normalised [N, embed_dim]."""
    x = sd[prefix + "token_embedding.weight"][text] + sd[prefix + "positional_embedding"]
    n, L, w = x.shape
    hd = w // heads
    mask = torch.full((L, L), float("-inf"), device=x.device).triu_(1)
    i = 0
    while f"{prefix}transformer.resblocks.{i}.ln_1.weight" in sd:
        p = f"{prefix}transformer.resblocks.{i}."
        y = F.layer_norm(x, (w,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
        qkv = F.linear(y, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
        q, k, v = [t.reshape(n, L, heads, hd).transpose(1, 2) for t in qkv.chunk(3, dim=-1)]
        att = torch.softmax(torch.matmul(q, k.transpose(2, 3)) * hd ** -0.5 + mask, dim=-1)
        y = torch.matmul(att, v).transpose(1, 2).reshape(n, L, w)
        x = x + F.linear(y, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
        y = F.layer_norm(x, (w,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
        y = F.gelu(F.linear(y, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]))
        x = x + F.linear(y, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
        i += 1
    x = F.layer_norm(x, (w,), sd[prefix + "ln_final.weight"], sd[prefix + "ln_final.bias"], 1e-5)
    return x[torch.arange(n, device=x.device), text.argmax(dim=-1)] @ sd[prefix + "text_projection"]


def degradation_logits(degra_features, text_features):
    """100 * cos-sim logits against the class prompts (evaluate_daclip.py:46-47,79-80)."""
    d = degra_features / degra_features.norm(dim=-1, keepdim=True)
    t = text_features / text_features.norm(dim=-1, keepdim=True)
    return 100.0 * d @ t.T


def degradation_argmax(degra_features, text_features):
    """argmax softmax(logits) (evaluate_daclip.py:80-81); int64 [B]."""
    return torch.argmax(torch.softmax(degradation_logits(degra_features, text_features), dim=-1), dim=-1)
