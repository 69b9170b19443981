"""DaCLIP with the reference's API (open_clip/daclip_model.py:17-76, transformer.py:288-555, model.py:187-249,
factory.py:365-404 of the reference): `encode_image(image, control=True)` -> (image_features, degra_features),
two ViT-B/32 towers, the control tower's per-layer hidden states (through their zero-init linears) added to the
frozen CLIP tower in REVERSED order (`control.pop()`, transformer.py:367-368); `encode_text(tokens)` -> the CLIP text
tower (12 causal pre-LN blocks over 77 tokens, end-of-text token pooled through ln_final and text_projection), used once
per deployment for the degradation prompts.

The module holds parameters under the reference's state-dict keys (`visual.*`, its alias `clip.visual.*`,
`visual_control.*` incl. `visual_control.transformer.zero_modules.N`, `logit_scale`, and the text side `clip.transformer.*`,
`clip.token_embedding.weight`, `clip.positional_embedding`, `clip.ln_final.*`, `clip.text_projection`, `clip.logit_scale`:
the 631 keys of a full checkpoint).

Execution: every GEMM (patch embedding, in_proj, out_proj, c_fc, c_proj, zero-linear) is the tcgen05 kernel with
tokens as pixels (M = 50*B); the residual stream stays fp32 (fp32 residual in the GEMM epilogue) so that the
degradation-type argmax is stable; LayerNorm / 50-token attention / pooling are small fused CUDA kernels.
"""
import os
from collections import OrderedDict

import torch
import torch.nn as nn

from . import lib as L
from . import ops

DISTORTIONS = ["motion-blurry", "hazy", "jpeg-compressed", "low-light", "noisy", "raindrop", "rainy", "shadowed",
               "snowy", "uncompleted"]                                         # options/test.yml:4


# ------------------------------------------------------------------------------------------------ parameter holders
class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise L.DacError("parameter holder: compute happens in the CUDA engine")


class _Mlp(_Holder):
    def __init__(self, w):
        super().__init__()
        self.c_fc = nn.Linear(w, 4 * w)
        self.gelu = nn.GELU()
        self.c_proj = nn.Linear(4 * w, w)


class _ResidualAttentionBlock(_Holder):  # transformer.py:189-244
    def __init__(self, w, heads):
        super().__init__()
        self.ln_1 = nn.LayerNorm(w)
        self.attn = nn.MultiheadAttention(w, heads)
        self.ln_2 = nn.LayerNorm(w)
        self.mlp = _Mlp(w)


class _Transformer(_Holder):  # transformer.py:328-369
    def __init__(self, w, layers, heads):
        super().__init__()
        self.width, self.layers = w, layers
        self.resblocks = nn.ModuleList([_ResidualAttentionBlock(w, heads) for _ in range(layers)])


class _ControlTransformer(_Holder):  # transformer.py:288-325
    def __init__(self, transformer):
        super().__init__()
        self.transformer = transformer
        self.zero_modules = nn.ModuleList([nn.Linear(transformer.width, transformer.width)
                                           for _ in range(transformer.layers)])
        for p in self.zero_modules.parameters():
            p.detach().zero_()


class _VisionTransformer(_Holder):  # transformer.py:372-555
    def __init__(self, image_size=224, patch=32, width=768, layers=12, heads=12, embed_dim=512, control=False):
        super().__init__()
        self.image_size, self.patch, self.width, self.layers, self.heads = image_size, patch, width, layers, heads
        self.output_dim = embed_dim
        g = image_size // patch
        scale = width ** -0.5
        self.conv1 = nn.Conv2d(3, width, patch, patch, bias=False)
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn(g * g + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        t = _Transformer(width, layers, heads)
        self.transformer = _ControlTransformer(t) if control else t
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, embed_dim))


class _TextSide(_Holder):
    """Text half of open_clip's CLIP (model.py:203-213) under the reference's `clip.*` keys; initialised like
    TextTransformer.init_parameters (transformer.py:609-627)."""

    def __init__(self, width=512, heads=8, layers=12, context_length=77, vocab_size=49408, embed_dim=512):
        super().__init__()
        self.width, self.heads, self.layers = width, heads, layers
        self.context_length, self.vocab_size = context_length, vocab_size
        self.transformer = _Transformer(width, layers, heads)
        self.token_embedding = nn.Embedding(vocab_size, width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, width))
        self.ln_final = nn.LayerNorm(width)
        self.text_projection = nn.Parameter(torch.empty(width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * 2.6592600)
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        proj_std, attn_std, fc_std = (width ** -0.5) * ((2 * layers) ** -0.5), width ** -0.5, (2 * width) ** -0.5
        for r in self.transformer.resblocks:
            nn.init.normal_(r.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(r.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(r.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(r.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=width ** -0.5)


class DaCLIP(nn.Module):
    """open_clip's `daclip_ViT-B-32` (model_configs/daclip_ViT-B-32.json; the defaults) and `daclip_ViT-L-14`
    (model_configs/daclip_ViT-L-14.json, the wild-ir encoder: see ARCHS)."""

    ARCHS = {"daclip_ViT-B-32": dict(image_size=224, patch=32, width=768, layers=12, heads=12, embed_dim=512,
                                     text_width=512, text_heads=8, text_layers=12),
             "daclip_ViT-L-14": dict(image_size=224, patch=14, width=1024, layers=24, heads=16, embed_dim=768,
                                     text_width=768, text_heads=12, text_layers=12)}

    def __init__(self, image_size=224, patch=32, width=768, layers=12, heads=12, embed_dim=512,
                 text_width=512, text_heads=8, text_layers=12, context_length=77, vocab_size=49408):
        super().__init__()
        self.visual = _VisionTransformer(image_size, patch, width, layers, heads, embed_dim)
        self.visual_control = _VisionTransformer(image_size, patch, width, layers, heads, embed_dim, control=True)
        self.logit_scale = nn.Parameter(torch.ones([]) * 2.6592600)
        self.clip = _TextSide(text_width, text_heads, text_layers, context_length, vocab_size, embed_dim)
        self.context_length, self.vocab_size = context_length, vocab_size
        self.has_text_weights = False         # set by load_reference_state_dict when the checkpoint carries the text side
        self._engines = {}
        self._text_engines = {}
        self._packed = None
        self._packed_text = None
        self.register_load_state_dict_post_hook(lambda m, keys: m.invalidate())

    def invalidate(self):
        self._engines, self._packed = {}, None
        self._text_engines, self._packed_text = {}, None

    @property
    def text_state(self):
        """The text-side tensors under their reference keys (`clip.*`)."""
        return OrderedDict((k, v) for k, v in self.state_dict().items() if k.startswith("clip."))

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)

    def load_reference_state_dict(self, sd):
        """Accepts the reference's checkpoint layout (factory.py:88-106): optional {'state_dict': ...} wrapper,
        optional 'module.' prefix, `clip.visual.*` alias of `visual.*`, text-tower keys kept aside."""
        if "state_dict" in sd and isinstance(sd["state_dict"], dict):
            sd = sd["state_dict"]
        own = self.state_dict()
        mine, unexpected = OrderedDict(), []
        for k, v in sd.items():
            k = k[7:] if k.startswith("module.") else k
            if k.startswith("clip.visual."):
                k = k[5:]
            if k in own:
                mine[k] = v
            else:
                unexpected.append(k)
        if unexpected:
            raise KeyError(f"checkpoint has {len(unexpected)} tensors this model does not know, e.g. {unexpected[:3]}")
        missing = [k for k in own if k not in mine and not k.startswith("clip.") and k != "logit_scale"]
        if missing:
            raise KeyError(f"checkpoint lacks {len(missing)} image-side tensors, e.g. {missing[:3]}")
        text_missing = [k for k in own if k.startswith("clip.") and k not in mine and k != "clip.logit_scale"]
        if text_missing and len(text_missing) != sum(k.startswith("clip.") and k != "clip.logit_scale" for k in own):
            raise KeyError(f"checkpoint has part of the text tower only; missing e.g. {text_missing[:3]}")
        self.has_text_weights = not text_missing        # image-side-only checkpoints keep the text side's init
        for k in own:
            mine.setdefault(k, own[k])
        self.load_state_dict(mine, strict=True)
        return self

    # ------------------------------------------------------------------ public API (daclip_model.py:46-55)
    def encode_image(self, image, control=False, normalize=False):
        """control=True: (image_features, degra_features) (daclip_model.py:47-52); control=False: the frozen CLIP tower
        alone, `self.clip.encode_image(image, normalize)` -> image_features (daclip_model.py:53-54, model.py:232-235)."""
        dev = self.visual.proj.device
        if dev.type != "cuda" or not image.is_cuda:
            raise L.DacError("DaCLIP (daclip_b200) runs on CUDA only (no CPU fallback)")
        B = image.shape[0]
        with L.on_device(dev):
            L.require_cuda(image)
            if self._packed is None:
                self._packed = _PackedDaCLIP(self)
            key = (B, bool(control))
            if key not in self._engines:
                self._engines[key] = _EncodeEngine(self._packed, self.visual, B, dev, control=bool(control))
            eng = self._engines[key]
            eng.image.copy_(image.to(torch.float32))
            eng.replay()
            img_f = eng.image_features.clone()
            if normalize:
                img_f = torch.nn.functional.normalize(img_f, dim=-1)      # plumbing on [B,512] outputs
            if not control:
                return img_f
            deg_f = eng.degra_features.clone()
            if normalize:
                deg_f = torch.nn.functional.normalize(deg_f, dim=-1)
            return img_f, deg_f

    def encode_text(self, text, normalize=False):
        """CLIP.encode_text (model.py:237-249) through DaCLIP.encode_text (daclip_model.py:55-56): text = int64 token ids
        [N, context_length] (see daclip_b200.tokenizer) -> fp32 [N, embed_dim] features of each row's end-of-text token."""
        dev = self.visual.proj.device
        if dev.type != "cuda":
            raise L.DacError("DaCLIP (daclip_b200) runs on CUDA only")
        if text.dim() != 2 or text.shape[1] != self.context_length:
            raise ValueError(f"encode_text takes [N, {self.context_length}] token ids, got {tuple(text.shape)}")
        if text.dtype not in (torch.int64, torch.int32):
            raise TypeError("encode_text takes integer token ids")
        if text.numel() and (int(text.min()) < 0 or int(text.max()) >= self.vocab_size):
            raise IndexError("index out of range in self")        # what nn.Embedding raises in the reference
        N = text.shape[0]
        with L.on_device(dev):
            if self._packed_text is None:
                self._packed_text = _PackedText(self.clip)
            if N not in self._text_engines:
                self._text_engines[N] = _TextEngine(self._packed_text, self.clip, N, dev)
            eng = self._text_engines[N]
            eng.tokens.copy_(text.to(device=dev, dtype=torch.int64))
            eng.replay()
            f = eng.features.clone()
        return torch.nn.functional.normalize(f, dim=-1) if normalize else f

    def degradation_argmax(self, degra_features, text_features, return_logits=False):
        """argmax_j softmax(100 * cos(degra, text_j)) (da-clip/src/evaluate_daclip.py:46-47,79-81)."""
        if not degra_features.is_cuda:
            raise L.DacError("degradation_argmax takes CUDA tensors only (no CPU fallback)")
        with L.on_device(degra_features):
            d = degra_features.to(torch.float32).contiguous()
            t = text_features.to(device=d.device, dtype=torch.float32).contiguous()
            logits = torch.empty(d.shape[0], t.shape[0], device=d.device)
            am = torch.empty(d.shape[0], dtype=torch.int64, device=d.device)
            ops.degradation_argmax(d, t, logits, am)
        return (am, logits) if return_logits else am


class _PackedTower:
    def __init__(self, vit, prefix_blocks, zero_modules=None):
        dev = vit.proj.device

        def f32(t):
            return t.detach().to(dev, torch.float32).contiguous()

        w1 = f32(vit.conv1.weight).reshape(vit.width, -1)
        kpad = -(-w1.shape[1] // 64) * 64
        self.conv1 = ops.pack_linear(torch.nn.functional.pad(w1, (0, kpad - w1.shape[1])))
        self.cls, self.pos = f32(vit.class_embedding), f32(vit.positional_embedding)
        self.ln_pre = (f32(vit.ln_pre.weight), f32(vit.ln_pre.bias))
        self.ln_post = (f32(vit.ln_post.weight), f32(vit.ln_post.bias))
        self.proj = f32(vit.proj)
        self.blocks = _pack_blocks(prefix_blocks, f32, zero_modules)


def _pack_blocks(resblocks, f32, zero_modules=None):
    blocks = []
    for i, r in enumerate(resblocks):
        blk = dict(
            ln1=(f32(r.ln_1.weight), f32(r.ln_1.bias)), ln2=(f32(r.ln_2.weight), f32(r.ln_2.bias)),
            qkv=ops.pack_linear(f32(r.attn.in_proj_weight)), qkv_b=f32(r.attn.in_proj_bias),
            out=ops.pack_linear(f32(r.attn.out_proj.weight)), out_b=f32(r.attn.out_proj.bias),
            fc=ops.pack_linear(f32(r.mlp.c_fc.weight)), fc_b=f32(r.mlp.c_fc.bias),
            proj=ops.pack_linear(f32(r.mlp.c_proj.weight)), proj_b=f32(r.mlp.c_proj.bias))
        if zero_modules is not None:
            blk["zero"] = ops.pack_linear(f32(zero_modules[i].weight))
            blk["zero_b"] = f32(zero_modules[i].bias)
        blocks.append(blk)
    return blocks


class _PackedText:
    def __init__(self, t: _TextSide):
        dev = t.text_projection.device

        def f32(x):
            return x.detach().to(dev, torch.float32).contiguous()

        self.emb, self.pos = f32(t.token_embedding.weight), f32(t.positional_embedding)
        self.ln_final = (f32(t.ln_final.weight), f32(t.ln_final.bias))
        self.proj = f32(t.text_projection)
        self.blocks = _pack_blocks(t.transformer.resblocks, f32)


class _PackedDaCLIP:
    def __init__(self, m: DaCLIP):
        self.clip = _PackedTower(m.visual, m.visual.transformer.resblocks)
        ct = m.visual_control.transformer
        self.control = _PackedTower(m.visual_control, ct.transformer.resblocks, ct.zero_modules)
        # CLIP layer i adds the hidden of control layer j = L-1-i through that layer's zero-linear (transformer.py:355-369:
        # control.pop()).  x + W_proj hid + Z_j xc_j is ONE GEMM over the virtual concat [hid | xc_j] against [W_proj | Z_j]:
        # no zero-linear launch, no hidden tensor written and read back, and the sum stays in fp32 (TMEM) until the stream
        Lb = len(self.clip.blocks)
        dev = m.visual.proj.device
        for i, blk in enumerate(self.clip.blocks):
            r, z = m.visual.transformer.resblocks[i], ct.zero_modules[Lb - 1 - i]
            wcat = torch.cat([r.mlp.c_proj.weight.detach().to(dev, torch.float32),
                              z.weight.detach().to(dev, torch.float32)], dim=1).contiguous()
            blk["proj_ctl"] = ops.pack_linear(wcat)
            blk["proj_ctl_b"] = (r.mlp.c_proj.bias.detach().to(dev, torch.float32)
                                 + z.bias.detach().to(dev, torch.float32)).contiguous()


class _EncodeEngine:
    """Launch plan of encode_image for a fixed batch, captured in one CUDA graph: control=True = the control tower
    followed by the CLIP tower that consumes its hidden states; control=False = the CLIP tower alone."""

    FUSE_ZERO = os.environ.get("DAC_FUSE_ZERO", "1") != "0"   # zero-linears inside the CLIP tower's c_proj GEMMs

    def __init__(self, pk: _PackedDaCLIP, vit, B, dev, control=True):
        self.B, self.dev = B, dev
        S, p, w, heads = vit.image_size, vit.patch, vit.width, vit.heads
        g = S // p
        Ltok = g * g + 1
        M = B * Ltok
        self.steps, self.flops = [], 0.0
        bf, f32 = dict(device=dev, dtype=torch.bfloat16), dict(device=dev, dtype=torch.float32)
        self.image = torch.zeros(B, 3, S, S, **f32)
        self.image_features = torch.zeros(B, vit.output_dim, **f32)
        self.degra_features = torch.zeros(B, vit.output_dim, **f32)
        kpad = -(-(3 * p * p) // 64) * 64           # GEMM K in 64-channel chunks (ViT-L/14: 588 -> 640, zero columns)
        patches = torch.zeros(1, 1, B * g * g, kpad, **bf)
        self.add(lambda: ops.vit_patchify(self.image, patches, B, S, p))
        hiddens = []

        def tower(tp, out_features, control_in=None, plain=False):
            pe = torch.zeros(1, 1, B * g * g, w, **bf)
            self.conv(patches, kpad, tp.conv1, pe, B * g * g)
            x = torch.zeros(1, 1, M, w, **f32)
            self.add(lambda: ops.vit_embed(pe, tp.cls, tp.pos, tp.ln_pre[0], tp.ln_pre[1], x, B, Ltok, w))
            n = torch.zeros(1, 1, M, w, **bf)
            qkv = torch.zeros(1, 1, M, 3 * w, **bf)
            att = torch.zeros(1, 1, M, w, **bf)
            hid = torch.zeros(1, 1, M, 4 * w, **bf)
            for i, blk in enumerate(tp.blocks):
                self.add(lambda blk=blk: ops.layernorm_rows_f32(x, n, M, w, blk["ln1"][0], blk["ln1"][1], 1e-5))
                self.conv(n, w, blk["qkv"], qkv, M, bias=blk["qkv_b"])
                self.add(lambda: ops.attention(qkv, att, B, Ltok, heads, w // heads))
                self.conv(att, w, blk["out"], None, M, bias=blk["out_b"], res_f32=x, out_f32=x)
                self.add(lambda blk=blk: ops.layernorm_rows_f32(x, n, M, w, blk["ln2"][0], blk["ln2"][1], 1e-5))
                self.conv(n, w, blk["fc"], hid, M, bias=blk["fc_b"], act=L.ACT_GELU)
                if plain:
                    self.conv(hid, 4 * w, blk["proj"], None, M, bias=blk["proj_b"], res_f32=x, out_f32=x)
                elif control_in is None:
                    # control tower: x <- x + mlp; also keep a bf16 copy as the zero-linear's GEMM operand
                    xb = torch.zeros(1, 1, M, w, **bf)
                    self.conv(hid, 4 * w, blk["proj"], xb, M, bias=blk["proj_b"], res_f32=x, out_f32=x)
                    if self.FUSE_ZERO:
                        hiddens.append(xb)              # the CLIP tower applies the zero-linear inside its c_proj GEMM
                    else:
                        h = torch.zeros(1, 1, M, w, **bf)
                        self.conv(xb, w, blk["zero"], h, M, bias=blk["zero_b"])
                        hiddens.append(h)
                elif self.FUSE_ZERO:
                    # CLIP tower: x <- x + [W_proj | Z_j] [hid | xc_j] + (b_proj + b_zero_j), j = L-1-i (control.pop())
                    self.conv(hid, 4 * w, blk["proj_ctl"], None, M, src1=control_in[len(tp.blocks) - 1 - i], c1=w,
                              bias=blk["proj_ctl_b"], res_f32=x, out_f32=x)
                else:
                    # CLIP tower: x <- x + mlp + control.pop()  (hidden of control layer L-1-i)
                    self.conv(hid, 4 * w, blk["proj"], None, M, bias=blk["proj_b"], res_f32=x, out_f32=x,
                              res2=control_in[len(tp.blocks) - 1 - i])
            self.add(lambda: ops.vit_pool(x, B, Ltok, w, tp.ln_post[0], tp.ln_post[1], tp.proj, out_features))

        if control:
            tower(pk.control, self.degra_features)
            tower(pk.clip, self.image_features, control_in=hiddens)
        else:
            tower(pk.clip, self.image_features, plain=True)
        self.flops += (2 if control else 1) * 4.0 * B * heads * Ltok * Ltok * (w // heads) * len(pk.clip.blocks)
        self.graph = None

    def add(self, fn):
        self.steps.append(fn)

    def conv(self, src, cin, pw, out, M, **kw):
        plan = ops.ConvPlan(src, cin, pw, out, B=1, H=1, W=M, **kw)
        self.flops += plan.flops
        self.steps.append(plan.run)

    def run_eager(self):
        for fn in self.steps:
            fn()

    def replay(self):
        if self.graph is None:
            self.run_eager()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.run_eager()
            self.graph = g
        self.graph.replay()


class _TextEngine(_EncodeEngine):
    """Launch plan of encode_text for a fixed number of prompts: the ViT block chain with causal attention, the
    residual stream in fp32, entry = embedding gather, exit = end-of-text pooling; one CUDA graph."""

    def __init__(self, pk: _PackedText, t: _TextSide, N, dev):
        self.B, self.dev = N, dev
        Ltok, w, heads = t.context_length, t.width, t.heads
        M = N * Ltok
        self.steps, self.flops = [], 0.0
        bf, f32 = dict(device=dev, dtype=torch.bfloat16), dict(device=dev, dtype=torch.float32)
        self.tokens = torch.zeros(N, Ltok, device=dev, dtype=torch.int64)
        self.eot = torch.zeros(N, device=dev, dtype=torch.int32)
        self.features = torch.zeros(N, pk.proj.shape[1], **f32)
        x = torch.zeros(1, 1, M, w, **f32)
        n = torch.zeros(1, 1, M, w, **bf)
        qkv = torch.zeros(1, 1, M, 3 * w, **bf)
        att = torch.zeros(1, 1, M, w, **bf)
        hid = torch.zeros(1, 1, M, 4 * w, **bf)
        self.add(lambda: ops.text_embed(self.tokens, pk.emb, pk.pos, x, self.eot, N, Ltok, w))
        for blk in pk.blocks:
            self.add(lambda blk=blk: ops.layernorm_rows_f32(x, n, M, w, blk["ln1"][0], blk["ln1"][1], 1e-5))
            self.conv(n, w, blk["qkv"], qkv, M, bias=blk["qkv_b"])
            self.add(lambda: ops.attention_causal(qkv, att, N, Ltok, heads, w // heads))
            self.conv(att, w, blk["out"], None, M, bias=blk["out_b"], res_f32=x, out_f32=x)
            self.add(lambda blk=blk: ops.layernorm_rows_f32(x, n, M, w, blk["ln2"][0], blk["ln2"][1], 1e-5))
            self.conv(n, w, blk["fc"], hid, M, bias=blk["fc_b"], act=L.ACT_GELU)
            self.conv(hid, 4 * w, blk["proj"], None, M, bias=blk["proj_b"], res_f32=x, out_f32=x)
        self.add(lambda: ops.text_pool(x, self.eot, N, Ltok, w, pk.ln_final[0], pk.ln_final[1], pk.proj, self.features))
        self.graph = None


def create_model_from_pretrained(model_name="daclip_ViT-B-32", pretrained=None, device="cuda", **_):
    """factory.py:365-404 for the one model on this path.  Returns (model, preprocess); `pretrained` is a path
    to a reference checkpoint (.pt: raw state dict or {'state_dict': ...}), or None for random init."""
    if model_name not in DaCLIP.ARCHS:
        raise NotImplementedError(f"{model_name}: only {sorted(DaCLIP.ARCHS)} are on the restoration path")
    model = DaCLIP(**DaCLIP.ARCHS[model_name])
    if pretrained:
        model.load_reference_state_dict(torch.load(pretrained, map_location="cpu"))
    return model.to(device).eval(), clip_preprocess


CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def clip_preprocess(pil_image, resolution=224):
    """data/util.py:87-93 / open_clip.transform: bicubic resize of the short side, centre crop, normalise.
    Host-side image IO (PIL/torchvision), outside the measured path."""
    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToTensor
    return Compose([Resize(resolution, interpolation=InterpolationMode.BICUBIC), CenterCrop(resolution), ToTensor(),
                    Normalize(CLIP_MEAN, CLIP_STD)])(pil_image)
