"""GPU: clip_transform / tensor2img kernels through the C ABI, bit-exact against the reference's own outputs
(tests/golden/imageio.pt) and against the numpy oracle on further sizes."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "imageio.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD)


def test_clip_transform_bit_exact_vs_reference_golden(cuda, gold):
    from daclip_b200 import imageio, synthetic
    mean = torch.tensor(imageio.CLIP_MEAN).view(3, 1, 1)
    std = torch.tensor(imageio.CLIP_STD).view(3, 1, 1)
    for g in gold["clip"]:
        img = synthetic.natural_image(g["h"], g["w"], seed=g["seed"])
        got = imageio.clip_transform(img).cpu()
        ref = (g["u8"].float() / 255 - mean) / std
        assert torch.equal(got, ref), f"{g['h']}x{g['w']}: {(got != ref).sum().item()} values differ"


@pytest.mark.parametrize("h,w", [(231, 224), (225, 640), (1000, 301), (64, 80), (224, 224), (2048, 1536)])
def test_clip_transform_bit_exact_vs_oracle(cuda, h, w):
    from daclip_b200 import imageio, synthetic
    from oracle import imageio_oracle as O
    img = synthetic.natural_image(h, w, seed=h + w)
    _, ref = O.clip_transform(img, imageio.pil_bicubic_coeffs, imageio.resized_size)
    got = imageio.clip_transform(torch.from_numpy(img).cuda()).cpu()
    assert torch.equal(got, torch.from_numpy(ref))


def test_clip_transform_batch_and_encoder_input_shape(cuda):
    from daclip_b200 import imageio, synthetic
    imgs = [synthetic.natural_image(h, w, seed=i) for i, (h, w) in enumerate([(256, 256), (300, 400), (512, 384)])]
    out = imageio.clip_transform_batch(imgs)
    assert out.shape == (3, 3, 224, 224) and out.dtype == torch.float32 and out.is_cuda
    for i, im in enumerate(imgs):
        assert torch.equal(out[i], imageio.clip_transform(im))
    with pytest.raises(ValueError):
        imageio.clip_transform(np.zeros((10, 10), np.float32))


def test_tensor2img_bit_exact_vs_reference_golden(cuda, gold):
    from daclip_b200 import imageio
    for g in gold["t2i"]:
        got = imageio.tensor2img(g["x"].cuda())
        ref = g["img"].numpy()
        assert got.dtype == np.uint8 and got.shape == ref.shape
        assert np.array_equal(got, ref), f"{tuple(g['x'].shape)}: {(got != ref).sum()} bytes differ"


def test_tensor2img_batch_and_ranges(cuda):
    from daclip_b200 import imageio
    from oracle import imageio_oracle as O
    g = torch.Generator().manual_seed(3)
    x = torch.rand(5, 3, 37, 53, generator=g) * 3 - 1
    got = imageio.tensor2img_batch(x.cuda(), min_max=(-1, 1)).cpu().numpy()
    for b in range(5):
        assert np.array_equal(got[b], O.tensor2img(x[b].numpy(), -1.0, 1.0))
    f = imageio.tensor2img(x[0].cuda(), out_type=np.float32)
    assert f.dtype == np.float32 and f.shape == (37, 53, 3) and 0.0 <= f.min() and f.max() <= 1.0
