"""Generates tests/golden/unet_wild.pt by running the REFERENCE's wild-ir ConditionalUNet
(config/wild-ir/models/modules/DenoisingUNet_arch.py: context_dim 768, use_degra_context false, scale 0.5 =
extra Downsample / Upsample pair) on seeded synthetic weights and inputs.  Build container only; outputs are committed.
TEST INFRASTRUCTURE - see oracle/__init__.py.  Separate from gen_golden.py because the two reference variants
share the module name `models`.

    python oracle/gen_golden_wild.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/universal-image-restoration"
sys.path.insert(0, os.path.join(REF, "config", "wild-ir"))
sys.path.insert(0, REF)
GOLD = os.path.join(ROOT, "tests", "golden")

from daclip_b200 import synthetic  # noqa: E402

WILD = dict(in_nc=3, out_nc=3, nf=64, ch_mult=[1, 2, 4, 8], context_dim=768, use_degra_context=False,
            use_image_context=True, scale=0.5)                                   # wild-ir/options/test.yml:31-41


def main():
    from models.modules.DenoisingUNet_arch import ConditionalUNet
    sd, kw = synthetic.unet_state_dict(5, **WILD)
    net = ConditionalUNet(**kw)
    net.load_state_dict(sd, strict=True)
    net.eval()
    out = {"ctor": kw, "weights_seed": 5, "cases": []}
    with torch.no_grad():
        for seed, (B, H, W), t in [(11, (2, 40, 24), 37.0), (12, (1, 64, 64), 100.0), (13, (1, 96, 64), 1.0)]:
            inp = synthetic.restoration_inputs(B, H, W, T=1, seed=seed, ctx_dim=768)
            xt = inp["lq"] + inp["eps0"] * (50 / 255)
            y = net(xt, inp["lq"], t, text_context=inp["text_context"], image_context=inp["image_context"])
            out["cases"].append(dict(seed=seed, shape=(B, H, W), time=t, out=y))
    torch.save(out, os.path.join(GOLD, "unet_wild.pt"))
    print("wrote", os.path.join(GOLD, "unet_wild.pt"))


if __name__ == "__main__":
    main()
