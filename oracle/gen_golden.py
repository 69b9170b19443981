"""Generates tests/golden/*.pt by running the REFERENCE modules (imported read-only from /root/reference) on
seeded synthetic weights and inputs.  Run in the build container only (the reference does not travel to the
GPU box); the outputs are committed.  TEST INFRASTRUCTURE - see oracle/__init__.py.

    python oracle/gen_golden.py

Shims: `ftfy` stub for the tokenizer import, no-op `.cuda()` while constructing
ControlTransformer on a CUDA-less host.  Per-step Gaussian noise is injected into the reference loops by
patching torch.randn_like, so both sides consume identical draws.
"""
import os
import sys
import types
import unittest.mock as um

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/universal-image-restoration"
sys.path.insert(0, os.path.join(REF, "config", "daclip-sde"))
sys.path.insert(0, REF)
GOLD = os.path.join(ROOT, "tests", "golden")

from daclip_b200 import synthetic  # noqa: E402

DISTORTIONS = ["motion-blurry", "hazy", "jpeg-compressed", "low-light", "noisy", "raindrop", "rainy", "shadowed",
               "snowy", "uncompleted"]          # options/test.yml:4


def gen_unet_and_sampler():
    from models.modules.DenoisingUNet_arch import ConditionalUNet
    from utils.sde_utils import IRSDE
    sd, kw = synthetic.unet_state_dict(0)
    net = ConditionalUNet(**kw)
    net.load_state_dict(sd, strict=True)
    net.eval()
    out = {"ctor": kw, "weights_seed": 0}
    with torch.no_grad():
        # (1) one forward on a size that needs reflect padding (40x24 -> 48x32)
        inp = synthetic.restoration_inputs(2, 40, 24, T=1, seed=1)
        xt = inp["lq"] + inp["eps0"] * (50 / 255)
        out["fwd_pad"] = dict(seed=1, shape=(2, 40, 24), time=37.0,
                              out=net(xt, inp["lq"], 37.0, text_context=inp["text_context"],
                                      image_context=inp["image_context"]))
        # (2) one forward at 64x64, t = 100 and t = 1
        inp = synthetic.restoration_inputs(1, 64, 64, T=1, seed=2)
        xt = inp["lq"] + inp["eps0"] * (50 / 255)
        out["fwd_64"] = dict(seed=2, shape=(1, 64, 64),
                             out_t100=net(xt, inp["lq"], 100.0, text_context=inp["text_context"],
                                          image_context=inp["image_context"]),
                             out_t1=net(xt, inp["lq"], 1.0, text_context=inp["text_context"],
                                        image_context=inp["image_context"]))
        # (3) full T=100 trajectories, both samplers + ODE, injected noise
        T = 100
        inp = synthetic.restoration_inputs(1, 32, 32, T=T, seed=3)
        sde = IRSDE(max_sigma=50, T=T, schedule="cosine", eps=0.005, device="cpu")
        sde.set_model(net)
        sde.set_mu(inp["lq"])
        x_T = inp["lq"] + inp["eps0"] * sde.max_sigma
        traj = dict(seed=3, shape=(1, 32, 32), T=T)
        for mode in ("sde", "posterior"):
            it = iter(inp["noise"])
            with um.patch("torch.randn_like", lambda t: next(it)):
                fn = sde.reverse_sde if mode == "sde" else sde.reverse_posterior
                traj[mode] = fn(x_T, text_context=inp["text_context"], image_context=inp["image_context"])
        # reverse_ode forwards no contexts (sde_utils.py:285), so with use_image_context=True the reference
        # itself raises a shape error inside attn2 (context=None -> self-attention through a 512-wide to_k):
        # the ODE loop is unreachable for this model; its per-step formula is pinned by oracle/sde_oracle.ode_step.
        out["trajectory"] = traj
        # schedule known answers
        out["schedule"] = dict(thetas=sde.thetas.clone(), sigmas=sde.sigmas.clone(),
                               thetas_cumsum=sde.thetas_cumsum.clone(), sigma_bars=sde.sigma_bars.clone(),
                               dt=sde.dt.clone())
    torch.save(out, os.path.join(GOLD, "unet_sampler.pt"))
    print("unet_sampler.pt written;  sde final absmax", traj["sde"].abs().max().item(),
          "posterior final absmax", traj["posterior"].abs().max().item())


def gen_reduced_step():
    """Reduced-step sampling: IRSDE(T=100, sample_T=20) - tables for 20 steps, network queried at 5 t."""
    from models.modules.DenoisingUNet_arch import ConditionalUNet
    from utils.sde_utils import IRSDE
    sd, kw = synthetic.unet_state_dict(0)
    net = ConditionalUNet(**kw)
    net.load_state_dict(sd, strict=True)
    net.eval()
    ST = 20
    inp = synthetic.restoration_inputs(1, 32, 32, T=ST, seed=8)
    sde = IRSDE(max_sigma=50, T=100, sample_T=ST, schedule="cosine", eps=0.005, device="cpu")
    sde.set_model(net)
    sde.set_mu(inp["lq"])
    x_T = inp["lq"] + inp["eps0"] * sde.max_sigma
    out = dict(seed=8, shape=(1, 32, 32), T=100, sample_T=ST, dt=sde.dt.clone(), sigma_bars=sde.sigma_bars.clone())
    with torch.no_grad():
        for mode in ("sde", "posterior"):
            it = iter(inp["noise"])
            with um.patch("torch.randn_like", lambda t: next(it)):
                fn = sde.reverse_sde if mode == "sde" else sde.reverse_posterior
                out[mode] = fn(x_T, text_context=inp["text_context"], image_context=inp["image_context"])
    torch.save(out, os.path.join(GOLD, "sampler_reduced.pt"))
    print("sampler_reduced.pt written", out["sde"].abs().max().item(), out["posterior"].abs().max().item())


def gen_trajectory_128():
    """Full T=100 reverse_sde / reverse_posterior of the reference at 128x128 (B = 1): every level is large enough
    for the production kernels (LinearAttention at 128^2 / 64^2 / 32^2 pixels, 256-token self-attention), unlike the
    32x32 trajectory of unet_sampler.pt."""
    from models.modules.DenoisingUNet_arch import ConditionalUNet
    from utils.sde_utils import IRSDE
    sd, kw = synthetic.unet_state_dict(0)
    net = ConditionalUNet(**kw)
    net.load_state_dict(sd, strict=True)
    net.eval()
    T = 100
    inp = synthetic.restoration_inputs(1, 128, 128, T=T, seed=13)
    sde = IRSDE(max_sigma=50, T=T, schedule="cosine", eps=0.005, device="cpu")
    sde.set_model(net)
    sde.set_mu(inp["lq"])
    x_T = inp["lq"] + inp["eps0"] * sde.max_sigma
    out = dict(seed=13, shape=(1, 128, 128), T=T, weights_seed=0)
    with torch.no_grad():
        for mode in ("sde", "posterior"):
            it = iter(inp["noise"])
            with um.patch("torch.randn_like", lambda t: next(it)):
                fn = sde.reverse_sde if mode == "sde" else sde.reverse_posterior
                out[mode] = fn(x_T, text_context=inp["text_context"], image_context=inp["image_context"])
    torch.save(out, os.path.join(GOLD, "trajectory_128.pt"))
    print("trajectory_128.pt written", out["sde"].abs().max().item(), out["posterior"].abs().max().item())


def gen_daclip():
    sys.modules.setdefault("ftfy", types.SimpleNamespace(fix_text=lambda s: s))
    import open_clip
    torch.manual_seed(20)
    with um.patch.object(torch.nn.Module, "cuda", lambda self, *a, **k: self):
        model = open_clip.create_model("daclip_ViT-B-32", pretrained=None, device="cpu")
    model.eval()
    vis = synthetic.daclip_visual_state_dict(10)
    full = model.state_dict()
    missing = [k for k in vis if k not in full]
    assert not missing, missing[:5]
    for k, v in vis.items():
        assert full[k].shape == v.shape, (k, full[k].shape, v.shape)
    model.load_state_dict(vis, strict=False)          # text tower keeps its seeded reference init
    # `visual.*` is aliased to `clip.visual.*` in the reference (same module object)
    assert model.visual is model.clip.visual
    tok = open_clip.get_tokenizer("ViT-B-32")
    g = torch.Generator().manual_seed(4)
    image = torch.randn(4, 3, 224, 224, generator=g)
    with torch.no_grad():
        text_features = model.encode_text(tok(DISTORTIONS))
        image_features, degra_features = model.encode_image(image, control=True)
        plain_features = model.encode_image(image, control=False)       # daclip_model.py:53-54: the CLIP tower alone
        d = degra_features / degra_features.norm(dim=-1, keepdim=True)
        t = text_features / text_features.norm(dim=-1, keepdim=True)
        probs = (100.0 * d @ t.T).softmax(dim=-1)       # evaluate_daclip.py:79-80
        pred = torch.argmax(probs, dim=-1)
    torch.save(dict(weights_seed=10, image_seed=4, text_features=text_features, image_features=image_features,
                    degra_features=degra_features, plain_features=plain_features, logits=100.0 * d @ t.T, argmax=pred),
               os.path.join(GOLD, "daclip.pt"))
    print("daclip.pt written; argmax", pred.tolist(), "top-2 gap",
          (probs.topk(2).values[:, 0] - probs.topk(2).values[:, 1]).tolist())


def gen_daclip_l14():
    """wild-ir encoder: daclip_ViT-L-14 (model_configs/daclip_ViT-L-14.json), both towers, B = 2."""
    sys.modules.setdefault("ftfy", types.SimpleNamespace(fix_text=lambda s: s))
    import open_clip
    torch.manual_seed(21)
    with um.patch.object(torch.nn.Module, "cuda", lambda self, *a, **k: self):
        model = open_clip.create_model("daclip_ViT-L-14", pretrained=None, device="cpu")
    model.eval()
    vis = synthetic.daclip_visual_state_dict(30, arch="ViT-L-14")
    full = model.state_dict()
    missing = [k for k in vis if k not in full]
    assert not missing, missing[:5]
    for k, v in vis.items():
        assert full[k].shape == v.shape, (k, full[k].shape, v.shape)
    model.load_state_dict(vis, strict=False)
    g = torch.Generator().manual_seed(6)
    image = torch.randn(2, 3, 224, 224, generator=g)
    with torch.no_grad():
        image_features, degra_features = model.encode_image(image, control=True)
    torch.save(dict(weights_seed=30, image_seed=6, image_features=image_features, degra_features=degra_features),
               os.path.join(GOLD, "daclip_l14.pt"))
    print("daclip_l14.pt written", image_features.shape, image_features.abs().mean().item())


TEXT_PROMPTS = DISTORTIONS + [
    "a photo of a rainy street at night, reflections on the wet asphalt",
    "Low-Light  indoor scene;   heavy JPEG artefacts &amp; motion blur!",
    "x",
    " ".join(f"token{i}" for i in range(60)),            # longer than the context: truncated, forced end-of-text
]


def gen_daclip_text():
    """CLIP.encode_text of the reference (open_clip/model.py:237-249) for both model configs, tokens from
    the reference tokenizer (open_clip/tokenizer.py:159-188), seeded synthetic text weights."""
    sys.modules.setdefault("ftfy", types.SimpleNamespace(fix_text=lambda s: s))
    import open_clip
    out = dict(prompts=TEXT_PROMPTS)
    for name, arch, seed in (("daclip_ViT-B-32", "ViT-B-32", 12), ("daclip_ViT-L-14", "ViT-L-14", 32)):
        torch.manual_seed(22)
        with um.patch.object(torch.nn.Module, "cuda", lambda self, *a, **k: self):
            model = open_clip.create_model(name, pretrained=None, device="cpu")
        model.eval()
        txt = synthetic.daclip_text_state_dict(seed, arch=arch)
        full = model.state_dict()
        for k, v in txt.items():
            assert full[k].shape == v.shape, (k, full[k].shape, v.shape)
        text_keys = [k for k in full if k.startswith("clip.") and not k.startswith("clip.visual.") and k != "clip.logit_scale"]
        assert sorted(text_keys) == sorted(txt), set(text_keys) ^ set(txt)
        model.load_state_dict(txt, strict=False)
        tokens = open_clip.get_tokenizer(name.replace("daclip_", ""))(TEXT_PROMPTS)
        with torch.no_grad():
            feats = model.encode_text(tokens)
        out["tokens"] = tokens.to(torch.int32)
        out[arch] = dict(weights_seed=seed, features=feats)
        print(name, feats.shape, feats.abs().mean().item(), tokens.argmax(-1).tolist())
    torch.save(out, os.path.join(GOLD, "daclip_text.pt"))


def gen_unet_noctx():
    """The class-default construction ConditionalUNet(use_image_context=False) (DenoisingUNet_arch.py:22-23): every level is
    a LinearAttention, incl. the 512-channel ones of level 3 / mid (:84-85,105).  Also reverse_ode, which forwards no
    contexts (sde_utils.py:282-294) and therefore only runs with this construction."""
    from models.modules.DenoisingUNet_arch import ConditionalUNet
    from utils.sde_utils import IRSDE
    ctor = dict(use_image_context=False)
    sd, kw = synthetic.unet_state_dict(5, **ctor)
    net = ConditionalUNet(**kw)
    net.load_state_dict(sd, strict=True)
    net.eval()
    cases = []
    with torch.no_grad():
        for seed, (B, H, W), time in ((21, (1, 64, 64), 100.0), (22, (2, 40, 24), 13.0)):
            inp = synthetic.restoration_inputs(B, H, W, T=1, seed=seed)
            xt = inp["lq"] + inp["eps0"] * (50 / 255)
            cases.append(dict(seed=seed, shape=(B, H, W), time=time,
                              out=net(xt, inp["lq"], time, text_context=inp["text_context"], image_context=None)))
        T = 100
        inp = synthetic.restoration_inputs(1, 32, 32, T=T, seed=23)
        sde = IRSDE(max_sigma=50, T=T, schedule="cosine", eps=0.005, device="cpu")
        sde.set_model(net)
        sde.set_mu(inp["lq"])
        x_T = inp["lq"] + inp["eps0"] * sde.max_sigma
        ode = dict(seed=23, shape=(1, 32, 32), T=T, out=sde.reverse_ode(x_T))
    torch.save(dict(ctor=ctor, weights_seed=5, cases=cases, ode=ode), os.path.join(GOLD, "unet_noctx.pt"))
    print("unet_noctx.pt written; absmax", [c["out"].abs().max().item() for c in cases], "ode absmax",
          ode["out"].abs().max().item())


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["unet", "daclip"]
    if "unet" in which:
        gen_unet_and_sampler()
    if "daclip" in which:
        gen_daclip()
    if "reduced" in which:
        gen_reduced_step()
    if "traj128" in which:
        gen_trajectory_128()
    if "daclip_l14" in which:
        gen_daclip_l14()
    if "daclip_text" in which:
        gen_daclip_text()
    if "unet_noctx" in which:
        gen_unet_noctx()
