"""Batched test driver: the per-image loop of the reference's config/daclip-sde/test.py:101-130 with
the batch parallelism the kernels are built for.  The reference hard-wires `batch_size=1, num_workers=0`
(data/__init__.py:30-33) and reads back only the first image of a batch (denoising_model.py:170); here images of
equal size are grouped into batches of up to `max_batch`, every step of the flow runs on the GPU (CLIP view,
contexts, noise_state, T-step loop, clamp / quantise / BGR), and only uint8 pixels cross PCIe.

Same objects and calls as the reference loop: `clip_model.encode_image(img4clip, control=True)`,
`sde.noise_state(LQ)`, `model.feed_data(...)`, `model.test(sde, mode=...)`, `util.tensor2img(...)`.
"""
from collections import OrderedDict

import numpy as np
import torch

from . import imageio


class BatchedRestorer:
    def __init__(self, model, sde, clip_model, sampling_mode="posterior", max_batch=16, text_features=None):
        """model: DenoisingModel (create_model(opt)); sde: IRSDE with set_model done; clip_model: DaCLIP;
        text_features: optional [classes, 512] encodings of the `distortion` prompts (options/test.yml:4) for the
        degradation-type argmax (da-clip/src/evaluate_daclip.py:77-84)."""
        self.model, self.sde, self.clip = model, sde, clip_model
        self.mode, self.max_batch, self.text_features = sampling_mode, int(max_batch), text_features
        self.device = model.device

    def _to_chw(self, img):
        """HWC float RGB [0,1] (numpy / tensor) -> CHW fp32 tensor on the device (LQGT_dataset.py:145-146)."""
        t = torch.from_numpy(np.ascontiguousarray(img, dtype=np.float32)) if isinstance(img, np.ndarray) else img.float()
        return t.to(self.device, non_blocking=True).permute(2, 0, 1).contiguous()

    def restore(self, lq_images, noise=None):
        """lq_images: list of HWC float32 RGB images in [0, 1].  Returns a list (input order) of dicts with
        `Output` (uint8 HWC BGR numpy, what test.py:129 saves), `Output_tensor` (fp32 CHW on the device, unclamped)
        and, when text_features were given, `degradation` (int class index).  `noise`: optional {(H, W): tensor
        [T, B, 3, H, W]} injected per-step noise per size group (parity runs)."""
        groups = OrderedDict()
        for i, im in enumerate(lq_images):
            groups.setdefault((im.shape[0], im.shape[1]), []).append(i)
        results = [None] * len(lq_images)
        for (h, w), idx in groups.items():
            for s in range(0, len(idx), self.max_batch):
                chunk = idx[s:s + self.max_batch]
                nz = None if noise is None else noise.get((h, w))
                if nz is not None:
                    if nz.shape[1] != len(idx):
                        raise ValueError(f"noise for size {(h, w)} has {nz.shape[1]} images, the group has {len(idx)}")
                    nz = nz[:, s:s + len(chunk)].contiguous()      # this chunk's rows of the group's [T, B, 3, H, W]
                self._restore_group([lq_images[i] for i in chunk], chunk, results, nz)
        return results

    @torch.no_grad()
    def _restore_group(self, imgs, idx, results, noise):
        lq = torch.stack([self._to_chw(im) for im in imgs])                       # [B, 3, H, W]
        dev_imgs = [t.permute(1, 2, 0) for t in lq]                               # HWC views already on the device
        img4clip = imageio.clip_transform_batch([v.contiguous() for v in dev_imgs])
        image_context, degra_context = self.clip.encode_image(img4clip, control=True)    # test.py:114-117
        image_context, degra_context = image_context.float(), degra_context.float()
        noisy_state = self.sde.noise_state(lq)                                    # test.py:119
        self.model.feed_data(noisy_state, lq, None, text_context=degra_context, image_context=image_context)
        if noise is not None:
            self.sde.set_mu(self.model.condition)
            self.model.output = (self.sde.reverse_sde if self.mode == "sde" else self.sde.reverse_posterior)(
                self.model.state, noise=noise, text_context=degra_context, image_context=image_context)
        else:
            self.model.test(self.sde, mode=self.mode, save_states=False)          # test.py:123
        out = self.model.output
        u8 = imageio.tensor2img_batch(out).cpu().numpy()                          # test.py:129, one D2H of bytes
        classes = None
        if self.text_features is not None:
            classes = self.clip.degradation_argmax(degra_context, self.text_features).cpu().tolist()
        for j, i in enumerate(idx):
            results[i] = dict(Output=u8[j], Output_tensor=out[j])
            if classes is not None:
                results[i]["degradation"] = classes[j]
