"""GPU pre / post-processing around the restoration path, same call shapes as the reference:

  clip_transform(np_image, resolution=224)   universal-image-restoration/data/util.py:87-93
  tensor2img(tensor, out_type, min_max)      universal-image-restoration/utils/img_utils.py:136-163

`clip_transform` reproduces torchvision's Resize(BICUBIC) on a PIL image bit for bit: the tap ranges and 22-bit
fixed-point coefficients of Pillow's two-pass 8-bit resampler are computed here on the host in float64 (a few
hundred numbers per image size, cached), the passes themselves run in da-clip_b200/csrc/imageio.cu.
"""
import ctypes as C
import functools
import math

import numpy as np
import torch

from . import lib as L

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)     # data/util.py:93
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
_PRECISION_BITS = 32 - 8 - 2                        # Pillow Resample.c


def _bicubic(x):
    """Pillow's bicubic_filter (a = -0.5), float64."""
    a = -0.5
    x = np.abs(x)
    return np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1,
                    np.where(x < 2.0, (((x - 5) * x + 8) * x - 4) * a, 0.0))


@functools.lru_cache(maxsize=64)
def pil_bicubic_coeffs(in_size, out_size):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the full-image box: (bounds int32 [out, 2] =
    {first tap, tap count}, kk int32 [out, ksize], ksize)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = _bicubic((np.arange(xmax, dtype=np.float64) + xmin - center + 0.5) * ss)
        ww = 0.0
        for v in w:                       # sequential sum, as the C loop
            ww += v
        if ww != 0.0:
            w = w / ww
        fx = w * float(1 << _PRECISION_BITS)
        kk[xx, :xmax] = np.where(w < 0, np.trunc(-0.5 + fx), np.trunc(0.5 + fx)).astype(np.int64)
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def resized_size(h, w, size):
    """torchvision Resize(int): short side -> size, long side -> int(size * long / short)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)       # (new_h, new_w)


@functools.lru_cache(maxsize=64)
def _device_coeffs(in_size, out_size, device):
    b, k, ksize = pil_bicubic_coeffs(in_size, out_size)
    return torch.from_numpy(b).to(device), torch.from_numpy(k).to(device), ksize


def clip_transform(image, resolution=224, out=None):
    """`image`: float32 RGB HWC in [0, 1] (numpy array or tensor; moved to the current CUDA device if needed).
    Returns the normalised fp32 [3, resolution, resolution] CLIP view on the device, equal bit for bit to the
    reference's PIL / torchvision pipeline."""
    lib = L.load()
    if isinstance(image, np.ndarray):
        image = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32))
    if image.dim() != 3 or image.shape[2] != 3:
        raise ValueError("clip_transform expects an HWC RGB image")
    if not image.is_cuda:
        image = image.cuda(non_blocking=True)
    image = image.contiguous().float()
    H, W = int(image.shape[0]), int(image.shape[1])
    nh, nw = resized_size(H, W, resolution)
    if nh < resolution or nw < resolution:
        raise ValueError("image too small for the centre crop")          # torchvision would zero-pad; not used here
    dev = image.device
    hb, hk, hks = _device_coeffs(W, nw, dev)
    vb, vk, vks = _device_coeffs(H, nh, dev)
    mid = torch.empty(H, nw, 3, device=dev, dtype=torch.uint8)
    if out is None:
        out = torch.empty(3, resolution, resolution, device=dev, dtype=torch.float32)
    L.require_cuda(out)
    top, left = int(round((nh - resolution) / 2.0)), int(round((nw - resolution) / 2.0))   # CenterCrop
    st = L.stream_ptr()
    L.check(lib.dac_clip_resample_h(image.data_ptr(), H, W, mid.data_ptr(), nw, hb.data_ptr(), hk.data_ptr(), hks,
                                    0, H, st))
    mean = (C.c_float * 3)(*CLIP_MEAN)
    std = (C.c_float * 3)(*CLIP_STD)
    L.check(lib.dac_clip_resample_v_norm(mid.data_ptr(), nw, 0, vb.data_ptr(), vk.data_ptr(), vks, top, left,
                                         resolution, mean, std, out.data_ptr(), st))
    return out


def clip_transform_batch(images, resolution=224):
    """A list of HWC images (any sizes) -> [B, 3, resolution, resolution] fp32 on the device."""
    out = None
    for i, im in enumerate(images):
        if out is None:
            dev = im.device if torch.is_tensor(im) and im.is_cuda else torch.device("cuda", torch.cuda.current_device())
            out = torch.empty(len(images), 3, resolution, resolution, device=dev, dtype=torch.float32)
        clip_transform(im, resolution, out=out[i])
    return out


def tensor2img_batch(tensor, min_max=(0, 1)):
    """fp32 [B, C, H, W] on the device -> uint8 [B, H, W, C] (BGR for C = 3) on the device."""
    lib = L.load()
    L.require_cuda(tensor)
    t = tensor.contiguous().float()
    B, Cn, H, W = t.shape
    out = torch.empty(B, H, W, Cn, device=t.device, dtype=torch.uint8)
    L.check(lib.dac_tensor2img(t.data_ptr(), out.data_ptr(), B, Cn, H, W, float(min_max[0]), float(min_max[1]),
                               L.stream_ptr()))
    return out


def tensor2img(tensor, out_type=np.uint8, min_max=(0, 1)):
    """Reference signature and result (HWC BGR uint8 numpy array for a 3-D tensor, HW for 2-D, a make_grid mosaic
    for 4-D); the clamp / scale / round / channel swap run on the device, only the bytes cross PCIe."""
    t = tensor.squeeze()
    if not t.is_cuda:
        t = t.cuda()
    if t.dim() == 4:
        from torchvision.utils import make_grid          # layout only (padding 2), as the reference
        t = make_grid(t.float().clamp(*min_max), nrow=int(math.sqrt(len(t))), normalize=False)
    if t.dim() == 3:
        x = t[None]
    elif t.dim() == 2:
        x = t[None, None]
    else:
        raise TypeError("Only support 4D, 3D and 2D tensor. But received with dimension: {:d}".format(t.dim()))
    if out_type != np.uint8:
        lo, hi = min_max
        y = (x.float().clamp(lo, hi) - lo) / (hi - lo)
        y = y[0].flip(0).permute(1, 2, 0) if t.dim() == 3 else y[0, 0]
        return y.cpu().numpy().astype(out_type)
    img = tensor2img_batch(x, min_max)[0]
    img = img if t.dim() == 3 else img[:, :, 0]
    return img.cpu().numpy()
