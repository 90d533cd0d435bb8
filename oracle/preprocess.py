"""Image preprocessing of the reference's test_step / validation_step, restated.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference: modules/lightning_modules/single.py:248-262 (and multi.py:89-103) - `transforms.Compose([Resize(384),
CenterCrop([384, 384]), ToTensor(), Normalize(mean, std)])` applied to the PIL image opened by data/dicom_id.py:91-92
(`Image.open(path).convert('RGB')`).  `reference_test_transforms` runs that torchvision pipeline itself (torchvision and
Pillow are installed); `pil_resize_bilinear` restates what it does to the pixels, which is Pillow's `ImagingResample`
(third-party, Pillow 12.2.0 src/libImaging/Resample.c: `precompute_coeffs`, `normalize_coeffs_8bpc`,
`ImagingResampleHorizontal_8bpc` / `Vertical_8bpc`): a separable, antialiased (support scaled by the shrink factor)
triangle filter evaluated in 22-bit fixed point with a uint8 intermediate image, horizontal pass first.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def resized_size(h: int, w: int, size: int = 384):
    """torchvision `_compute_resized_output_size` for an int size: the shorter edge becomes `size`."""
    if w <= h:
        return int(size * h / w), size
    return size, int(size * w / h)


def center_crop_offsets(h: int, w: int, size: int = 384):
    """torchvision `center_crop`: int(round((h - size) / 2.0)) (Python's round-half-even)."""
    return int(round((h - size) / 2.0)), int(round((w - size) / 2.0))


def coeffs(in_size: int, out_size: int):
    """(bounds [out, 2] int32 (first tap, taps), kk [out, ksize] int32 fixed-point weights) of one axis."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.zeros(ksize, np.float64)
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
        tot = 0.0                # Pillow accumulates the weights in tap order
        for x in range(xmax):
            tot += w[x]
        if tot != 0.0:
            w[:xmax] /= tot
        for x in range(ksize):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """img uint8 [H, W, C]; one pass of the 8bpc resampler along `axis` (0 = vertical, 1 = horizontal)."""
    bounds, kk = coeffs(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.zeros((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        x0, n = bounds[xx]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[xx, :n].astype(np.int64), src[x0:x0 + n], axes=(0, 0))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_bilinear(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """uint8 [H, W, C] -> uint8 [out_h, out_w, C], bit-exact with PIL.Image.resize((out_w, out_h), BILINEAR)."""
    if img.shape[1] != out_w:
        img = _resample_axis(img, out_w, 1)
    if img.shape[0] != out_h:
        img = _resample_axis(img, out_h, 0)
    return img


def test_transforms_restated(img: np.ndarray, mean, std, size: int = 384) -> np.ndarray:
    """uint8 [H, W] or [H, W, 3] -> float32 [3, size, size]: convert('RGB') + the reference's test_transforms."""
    if img.ndim == 2:
        img = np.repeat(img[:, :, None], 3, axis=2)
    h, w = img.shape[:2]
    nh, nw = resized_size(h, w, size)
    r = pil_resize_bilinear(img, nh, nw)
    top, left = center_crop_offsets(nh, nw, size)
    c = r[top:top + size, left:left + size].astype(np.float32) / np.float32(255.0)
    c = (c - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)
    return np.ascontiguousarray(c.transpose(2, 0, 1))


def reference_test_transforms(img: np.ndarray, mean, std, size: int = 384) -> np.ndarray:
    """the reference's own pipeline (single.py:248-262) on the PIL image (data/dicom_id.py:91-92)"""
    from PIL import Image
    from torchvision import transforms
    t = transforms.Compose([transforms.Resize(size=size), transforms.CenterCrop(size=[size, size]), transforms.ToTensor(),
                            transforms.Normalize(mean=mean, std=std)])
    return t(Image.fromarray(img).convert("RGB")).numpy()
