"""CPU restatement (numpy, integer arithmetic) of the pre / post-processing around the path.  TEST INFRASTRUCTURE -
see oracle/__init__.py: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.

The arithmetic of `clip_transform` (universal-image-restoration/data/util.py:87-93) lives in third-party code that
is not under /root/reference: Pillow's ImagingResample (src/libImaging/Resample.c; the reference pins
`Pillow==9.5.0`, this image has 12.2.0 - the 8-bit resampler is unchanged between them) reached through torchvision's
Resize / CenterCrop / ToTensor / Normalize.  Restated here from its published algorithm: per output pixel a tap
window [xmin, xmin + n) around centre (xx + 0.5) * scale with half-width 2 * max(scale, 1), bicubic (a = -0.5)
weights normalised in float64, rounded to 22-bit fixed point, horizontal pass -> uint8 -> vertical pass -> uint8.
Pinned by tests/test_imageio_cpu.py against outputs of the reference function itself (tests/golden/imageio.pt,
made by oracle/gen_golden_imageio.py).  The coefficient tables are taken from the product's host code
(daclip_b200.imageio.pil_bicubic_coeffs) - they are what is being pinned.
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resample_axis0(u8, bounds, kk):
    """u8 [N, ...] uint8 resampled along axis 0 with Pillow's fixed-point accumulation (int32 wrap-around kept)."""
    out = np.empty((bounds.shape[0],) + u8.shape[1:], np.uint8)
    for i, (lo, n) in enumerate(bounds):
        acc = np.full(u8.shape[1:], 1 << (PRECISION_BITS - 1), np.int32)
        for j in range(n):
            acc = acc + u8[lo + j].astype(np.int32) * np.int32(kk[i, j])
        out[i] = _clip8(acc)
    return out


def clip_transform(img, coeffs, resized_size, resolution=224,
                   mean=(0.48145466, 0.4578275, 0.40821073), std=(0.26862954, 0.26130258, 0.27577711)):
    """img float32 HWC RGB in [0, 1] -> (uint8 crop [res, res, 3], normalised float32 [3, res, res])."""
    u8 = (img * np.float32(255)).astype(np.uint8)                       # util.py:88
    h, w = u8.shape[:2]
    nh, nw = resized_size(h, w, resolution)
    hb, hk, _ = coeffs(w, nw)
    vb, vk, _ = coeffs(h, nh)
    mid = resample_axis0(np.ascontiguousarray(u8.transpose(1, 0, 2)), hb, hk).transpose(1, 0, 2)   # horizontal first
    res = resample_axis0(np.ascontiguousarray(mid), vb, vk)
    top, left = int(round((nh - resolution) / 2.0)), int(round((nw - resolution) / 2.0))
    crop = res[top:top + resolution, left:left + resolution]
    x = crop.astype(np.float32) / np.float32(255)
    x = (x - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)
    return crop, np.ascontiguousarray(x.transpose(2, 0, 1))


def tensor2img(x, lo=0.0, hi=1.0):
    """utils/img_utils.py:136-163 for a [C, H, W] float32 array: uint8 HWC, channels reversed."""
    t = np.clip(x.astype(np.float32), np.float32(lo), np.float32(hi))
    t = (t - np.float32(lo)) / np.float32(hi - lo)
    img = np.transpose(t[::-1], (1, 2, 0))
    return np.round(img * np.float32(255.0)).astype(np.uint8)
