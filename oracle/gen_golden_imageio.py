"""Generates tests/golden/imageio.pt by running the REFERENCE's own `clip_transform` (data/util.py:87-93, PIL +
torchvision) and `tensor2img` (utils/img_utils.py:136-163) on seeded synthetic images.  Build container only; the
outputs are committed.  TEST INFRASTRUCTURE - see oracle/__init__.py.

    python oracle/gen_golden_imageio.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/universal-image-restoration"
sys.path.insert(0, REF)
GOLD = os.path.join(ROOT, "tests", "golden")

from daclip_b200 import synthetic  # noqa: E402

SIZES = [(480, 720), (97, 141), (256, 256), (300, 224), (224, 513)]


def main():
    from data.util import clip_transform
    from utils.img_utils import tensor2img
    out = {"clip": [], "t2i": []}
    for i, (h, w) in enumerate(SIZES):
        img = synthetic.natural_image(h, w, seed=40 + i)                       # float32 HWC RGB in [0, 1]
        ref = clip_transform(img)                                              # [3, 224, 224] fp32
        mean = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(3, 1, 1)
        std = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(3, 1, 1)
        u8 = torch.round((ref * std + mean) * 255).to(torch.uint8)             # the cropped uint8 pixels
        assert torch.equal(((u8.float() / 255) - mean) / std, ref), "uint8 round trip of the golden is not exact"
        out["clip"].append(dict(seed=40 + i, h=h, w=w, u8=u8))
    g = torch.Generator().manual_seed(7)
    for shape in [(3, 64, 48), (1, 3, 33, 57), (40, 24), (4, 3, 16, 20)]:
        t = torch.rand(*shape, generator=g) * 1.4 - 0.2
        t.view(-1)[:8] = torch.tensor([0.5 / 255, 1.5 / 255, 2.5 / 255, -1.0, 2.0, 0.0, 1.0, 254.5 / 255])
        out["t2i"].append(dict(x=t.clone(), img=torch.from_numpy(tensor2img(t.clone()))))
    torch.save(out, os.path.join(GOLD, "imageio.pt"))
    print("wrote", os.path.join(GOLD, "imageio.pt"), os.path.getsize(os.path.join(GOLD, "imageio.pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
