"""GPU replacement of the reference's `test_transforms` (modules/lightning_modules/single.py:248-262, multi.py:89-103):
Resize(384) -> CenterCrop([384, 384]) -> ToTensor -> Normalize(mean, std) on the decoded image
(`Image.open(path).convert('RGB')`, data/dicom_id.py:91-92), through cxrm_preprocess_image.

    tf = TestTransforms(engine, mean, std)            # mean / std of the checkpoint's image processor (single.py:226)
    pixels = tf.batch([[img_a, img_b], [img_c]], max_images=5)   # -> [B, N, 3, 384, 384] fp32 on the device, zero padded
    enc = model.encoder(pixels)

Images are uint8 arrays / tensors [H, W] (grey, broadcast to three channels like convert('RGB')) or [H, W, 3], on the
host or already on the device.  JPEG decoding stays on the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from .engine import Engine, _stream

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class TestTransforms:
    __test__ = False     # not a pytest class

    def __init__(self, engine: Engine, mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD,
                 size: int = 384):
        assert len(mean) == 3 and len(std) == 3
        self.engine, self.size = engine, size
        self._mean = (C.c_float * 3)(*mean)
        self._std = (C.c_float * 3)(*std)

    def __call__(self, img, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """one image -> [3, size, size] fp32 on the engine's device (written into `out` when given)"""
        t = torch.as_tensor(img)
        if t.dtype != torch.uint8 or t.dim() not in (2, 3) or (t.dim() == 3 and t.shape[2] != 3):
            raise ValueError("image must be uint8 [H, W] or [H, W, 3]")
        t = t.contiguous()
        H, W = t.shape[:2]
        ch = 1 if t.dim() == 2 else 3
        if out is None:
            out = torch.empty(3, self.size, self.size, dtype=torch.float32, device=self.engine.device)
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.shape == (3, self.size, self.size)
        e = self.engine
        e._check(e.lib.cxrm_preprocess_image(e.h, C.c_void_p(t.data_ptr()), H, W, ch, W * ch, int(t.is_cuda), self.size,
                                             self._mean, self._std, C.c_void_p(out.data_ptr()), _stream()),
                 "cxrm_preprocess_image")
        return out

    def batch(self, studies, max_images: Optional[int] = None) -> torch.Tensor:
        """list (studies) of lists (images) -> [B, N, 3, size, size], missing image slots exactly zero (the reference
        pads studies with zero images and detects them by `pixel[b, n, 0, 0, 0] == 0`, modelling_longitudinal.py:83)."""
        N = max_images or max(len(s) for s in studies)
        px = torch.zeros(len(studies), N, 3, self.size, self.size, dtype=torch.float32, device=self.engine.device)
        for b, imgs in enumerate(studies):
            if len(imgs) > N:
                raise ValueError(f"study {b} has {len(imgs)} images, more than max_images={N}")
            for n, img in enumerate(imgs):
                self(img, out=px[b, n])
        return px
