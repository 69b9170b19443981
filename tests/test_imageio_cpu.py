"""CPU: the host side of clip_transform (Pillow coefficient tables, torchvision size rules) and the numpy oracle,
pinned against outputs of the reference's own functions (tests/golden/imageio.pt)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "imageio.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD)


def test_oracle_clip_transform_matches_reference_golden(gold):
    from daclip_b200 import imageio, synthetic
    from oracle import imageio_oracle as O
    for g in gold["clip"]:
        img = synthetic.natural_image(g["h"], g["w"], seed=g["seed"])
        crop, x = O.clip_transform(img, imageio.pil_bicubic_coeffs, imageio.resized_size)
        ref = g["u8"].permute(1, 2, 0).numpy()
        assert np.array_equal(crop, ref), f"{g['h']}x{g['w']}: {(crop != ref).sum()} bytes differ"
        mean = torch.tensor(imageio.CLIP_MEAN).view(3, 1, 1)
        std = torch.tensor(imageio.CLIP_STD).view(3, 1, 1)
        assert torch.equal(torch.from_numpy(x), (g["u8"].float() / 255 - mean) / std)


def test_oracle_clip_transform_matches_live_pillow():
    """Same check against Pillow / torchvision themselves when they are importable (more sizes, incl. upscaling)."""
    pytest.importorskip("PIL")
    tvt = pytest.importorskip("torchvision.transforms")
    from PIL import Image
    from daclip_b200 import imageio, synthetic
    from oracle import imageio_oracle as O
    tf = tvt.Compose([tvt.Resize(224, interpolation=tvt.InterpolationMode.BICUBIC), tvt.CenterCrop(224)])
    for i, (h, w) in enumerate([(231, 224), (225, 640), (1000, 301), (64, 80), (224, 224)]):
        img = synthetic.natural_image(h, w, seed=90 + i)
        ref = np.asarray(tf(Image.fromarray((img * 255).astype(np.uint8))))
        crop, _ = O.clip_transform(img, imageio.pil_bicubic_coeffs, imageio.resized_size)
        assert np.array_equal(crop, ref), f"{h}x{w}: {(crop != ref).sum()} bytes differ"


def test_resized_size_rules():
    from daclip_b200 import imageio
    assert imageio.resized_size(480, 720, 224) == (224, 336)
    assert imageio.resized_size(720, 480, 224) == (336, 224)
    assert imageio.resized_size(97, 141, 224) == (224, 325)
    assert imageio.resized_size(256, 256, 224) == (224, 224)


def test_oracle_tensor2img_matches_reference_golden(gold):
    from oracle import imageio_oracle as O
    for g in gold["t2i"]:
        x = g["x"].squeeze()
        if x.dim() == 4:
            continue                                    # mosaic layout: covered on the GPU side
        arr = x.numpy() if x.dim() == 3 else x.numpy()[None]
        got = O.tensor2img(arr)
        got = got if x.dim() == 3 else got[:, :, 0]
        assert np.array_equal(got, g["img"].numpy())
