"""GPU parity of DaCLIP.encode_image(control=True) and the degradation-type argmax against outputs of the
REFERENCE model (tests/golden/daclip.pt) and against the fp32 oracle at a larger batch.
Tolerance: features within 2 % of their range (bf16 GEMM operands, fp32 residual stream / LayerNorm / softmax);
the argmax is bit-exact (north_star)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def setup(cuda):
    from daclip_b200 import synthetic
    from daclip_b200.daclip import DaCLIP
    g = torch.load(os.path.join(GOLD, "daclip.pt"), weights_only=False)
    sd = synthetic.daclip_visual_state_dict(g["weights_seed"])
    m = DaCLIP().load_reference_state_dict(sd).to(cuda).eval()
    return m, sd, g


def rel(a, b):
    return (a - b).abs().max().item() / b.abs().max().item()


def test_encode_image_vs_reference_golden(setup):
    m, sd, g = setup
    image = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(g["image_seed"])).cuda()
    img_f, deg_f = m.encode_image(image, control=True)
    assert img_f.shape == (4, 512) and deg_f.dtype == torch.float32
    assert rel(img_f.cpu(), g["image_features"]) < 2e-2, rel(img_f.cpu(), g["image_features"])
    assert rel(deg_f.cpu(), g["degra_features"]) < 2e-2, rel(deg_f.cpu(), g["degra_features"])
    am, logits = m.degradation_argmax(deg_f, g["text_features"].cuda(), return_logits=True)
    assert torch.equal(am.cpu(), g["argmax"]), (am.cpu(), g["argmax"], (logits.cpu() - g["logits"]).abs().max())
    assert (logits.cpu() - g["logits"]).abs().max().item() < 0.05


def test_reference_checkpoint_layout_roundtrip(setup):
    """A full-checkpoint-style dict: {'state_dict': {module.clip.visual.*, module.visual_control.*, text keys}}."""
    from daclip_b200.daclip import DaCLIP
    m, sd, g = setup
    full = {}
    for k, v in sd.items():
        if k.startswith("visual."):
            full["module.clip." + k] = v
            full["module." + k] = v
        else:
            full["module." + k] = v
    from daclip_b200 import synthetic
    txt = synthetic.daclip_text_state_dict(12)
    full.update({"module." + k: v for k, v in txt.items()})
    full["module.clip.logit_scale"] = torch.tensor(1.5)
    m2 = DaCLIP().load_reference_state_dict({"state_dict": full})
    assert m2.has_text_weights and torch.equal(m2.text_state["clip.token_embedding.weight"], txt["clip.token_embedding.weight"])
    assert len(m2.state_dict()) + sum(k.startswith("visual.") for k in m2.state_dict()) == 631   # + the clip.visual.* alias
    with pytest.raises(KeyError):                   # a partial text tower is an error, not a silent random init
        DaCLIP().load_reference_state_dict({k: v for k, v in full.items() if "ln_final" not in k})
    for k, v in m.state_dict().items():
        if not k.startswith("clip."):               # `m` was loaded from an image-side-only checkpoint
            assert torch.equal(m2.state_dict()[k].cpu(), v.cpu()), k


def test_encode_image_batch_vs_oracle(setup):
    from oracle import daclip_oracle as D
    m, sd, g = setup
    B = 16
    image = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(11)).cuda()
    sdc = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref_img, ref_deg = D.encode_image_control(sdc, image)
    img_f, deg_f = m.encode_image(image, control=True)
    assert rel(img_f, ref_img) < 2e-2 and rel(deg_f, ref_deg) < 2e-2
    text = g["text_features"].cuda()
    am, logits = m.degradation_argmax(deg_f, text, return_logits=True)
    ref_logits = D.degradation_logits(ref_deg, text)
    top2 = ref_logits.topk(2, dim=-1).values
    gap = (top2[:, 0] - top2[:, 1])
    # bit-exact wherever the reference's own top-2 gap exceeds the measured logit error; report the gap otherwise
    err = (logits - ref_logits).abs().max().item()
    decided = gap > 4 * err
    assert torch.equal(am[decided], D.degradation_argmax(ref_deg, text)[decided])
    assert decided.float().mean().item() > 0.7, (gap, err)
    # normalize=True path
    a, b = m.encode_image(image, control=True, normalize=True)
    assert (a.norm(dim=-1) - 1).abs().max().item() < 1e-4 and (b.norm(dim=-1) - 1).abs().max().item() < 1e-4


def test_encode_image_batch256_argmax(setup):
    """BASELINE config C5: encode_image alone at batch 256 (M = 12,800 tokens), degradation-type argmax."""
    from oracle import daclip_oracle as D
    m, sd, g = setup
    B = 256
    image = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(21)).cuda()
    img_f, deg_f = m.encode_image(image, control=True)
    text = g["text_features"].cuda()
    am, logits = m.degradation_argmax(deg_f, text, return_logits=True)
    sdc = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        refs = [D.encode_image_control(sdc, image[i:i + 32]) for i in range(0, B, 32)]
    ref_img, ref_deg = torch.cat([r[0] for r in refs]), torch.cat([r[1] for r in refs])
    assert rel(img_f, ref_img) < 2e-2 and rel(deg_f, ref_deg) < 2e-2
    ref_logits = D.degradation_logits(ref_deg, text)
    err = (logits - ref_logits).abs().max().item()
    top2 = ref_logits.topk(2, dim=-1).values
    decided = (top2[:, 0] - top2[:, 1]) > 4 * err
    assert torch.equal(am[decided], ref_logits.argmax(-1)[decided])
    assert decided.float().mean().item() > 0.7


def test_vit_l14_encode_vs_reference_golden(cuda):
    """the wild-ir encoder daclip_ViT-L-14 (patch 14 -> K = 588 padded to 640, 257 tokens, d = 64 heads,
    24 + 24 layers) against features of the reference model itself."""
    from daclip_b200 import synthetic
    from daclip_b200.daclip import create_model_from_pretrained
    g = torch.load(os.path.join(GOLD, "daclip_l14.pt"), weights_only=False)
    m, _ = create_model_from_pretrained("daclip_ViT-L-14", device=cuda)
    m.load_reference_state_dict(synthetic.daclip_visual_state_dict(g["weights_seed"], arch="ViT-L-14"))
    m = m.to(cuda).eval()
    image = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(g["image_seed"])).cuda()
    img_f, deg_f = m.encode_image(image, control=True)
    assert img_f.shape == (2, 768) and deg_f.shape == (2, 768)
    assert rel(img_f.cpu(), g["image_features"]) < 2e-2, rel(img_f.cpu(), g["image_features"])
    assert rel(deg_f.cpu(), g["degra_features"]) < 2e-2, rel(deg_f.cpu(), g["degra_features"])


@pytest.mark.parametrize("name,arch", [("daclip_ViT-B-32", "ViT-B-32"), ("daclip_ViT-L-14", "ViT-L-14")])
def test_encode_text_vs_reference_golden(cuda, name, arch):
    """DaCLIP.encode_text (embedding gather, 12 causal blocks on the tcgen05 GEMM + causal flash attention,
    end-of-text pooling) against CLIP.encode_text of the reference on the reference tokenizer's ids."""
    from daclip_b200 import synthetic
    from daclip_b200.daclip import DaCLIP
    g = torch.load(os.path.join(GOLD, "daclip_text.pt"), weights_only=False)
    sd = synthetic.daclip_visual_state_dict(10 if arch == "ViT-B-32" else 30, arch=arch)
    sd.update(synthetic.daclip_text_state_dict(g[arch]["weights_seed"], arch=arch))
    m = DaCLIP(**DaCLIP.ARCHS[name]).load_reference_state_dict(sd).to(cuda).eval()
    assert m.has_text_weights
    tokens = g["tokens"].long()
    f = m.encode_text(tokens)                       # ids on the host, as the tokenizer returns them
    ref = g[arch]["features"]
    assert f.shape == ref.shape and f.dtype == torch.float32
    assert rel(f.cpu(), ref) < 2e-2, rel(f.cpu(), ref)
    # replay with ids already on the device, a different batch, normalised output
    f2 = m.encode_text(tokens[3:8].cuda(), normalize=True)
    want = torch.nn.functional.normalize(ref[3:8], dim=-1)
    assert (f2.cpu() - want).abs().max().item() < 2e-2 * want.abs().max().item()
    assert torch.equal(m.encode_text(tokens), f)    # graph replay is deterministic
    with pytest.raises(IndexError):
        m.encode_text(torch.full((1, 77), 49408))
    with pytest.raises(ValueError):
        m.encode_text(tokens[:, :40])


def test_text_tower_feeds_degradation_argmax(setup):
    """End of the DA-CLIP flow (evaluate_daclip.py:42-50,77-84) with text features from this encode_text instead of a
    constant tensor: logits against the oracle's on the same weights."""
    from daclip_b200 import synthetic
    from daclip_b200.daclip import DaCLIP
    from oracle import daclip_oracle as D
    m0, sd, g = setup
    gt = torch.load(os.path.join(GOLD, "daclip_text.pt"), weights_only=False)
    full = dict(sd)
    txt = synthetic.daclip_text_state_dict(gt["ViT-B-32"]["weights_seed"])
    full.update(txt)
    m = DaCLIP().load_reference_state_dict(full).cuda().eval()
    tokens = gt["tokens"][:10].long()
    text_f = m.encode_text(tokens)
    image = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(g["image_seed"])).cuda()
    _, deg_f = m.encode_image(image, control=True)
    am, logits = m.degradation_argmax(deg_f, text_f, return_logits=True)
    with torch.no_grad():
        ref_logits = D.degradation_logits(g["degra_features"], D.encode_text(txt, tokens))
    err = (logits.cpu() - ref_logits).abs().max().item()
    assert err < 0.6, err                           # logits are 100 * cosine: 0.6 = 6e-3 in cosine
    top2 = ref_logits.topk(2, dim=-1).values
    decided = (top2[:, 0] - top2[:, 1]) > 4 * err
    assert torch.equal(am.cpu()[decided], ref_logits.argmax(-1)[decided])
