"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's frame preprocessing.

Reference: ``image_transform`` EVA_clip/eva_clip.py:120-153 = torchvision ``Resize(224, BICUBIC)`` → ``CenterCrop(224)`` →
``_convert_to_rgb`` → ``ToTensor()`` → ``Normalize(mean, std)``, applied to PIL images at
inference_video_retrieval.py:43-49 and extract_features.py:48-50.

The arithmetic lives in two third-party dependencies that are not under /root/reference (requirements.txt: ``torchvision``,
``Pillow``, both unpinned; this container has torchvision 0.26.0 and Pillow 12.2.0):

* torchvision ``Resize(int)``: shorter edge → ``size``, longer edge → ``int(size * long / short)``; ``CenterCrop``: offsets
  ``int(round((dim - size) / 2.0))`` (Python banker's rounding).
* Pillow ``Image.resize(..., BICUBIC)`` on 8-bit images (``ImagingResample``, 8bpc path): separable, horizontal pass
  first into an 8-bit intermediate, then vertical; per output pixel the filter window is ``[center - support,
  center + support)`` with ``support = 2 * max(scale, 1)``, bicubic kernel with a = -0.5 evaluated in double, weights
  normalised to sum 1, converted to fixed point with 22 fractional bits (round half away from zero), accumulated in
  int32 starting from ``1 << 21`` and shifted right by 22 with saturation to [0, 255].

This is restated below in numpy (int64 accumulation of the same integer products) and pinned against PIL + torchvision
run in the build container: ``oracle/make_golden_preprocess.py`` → ``tests/golden/preprocess.npz``.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Pillow ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for the full-image box.
    Returns (bounds [out,2] int32 = (xmin, count), coeffs [out, ksize] int32, ksize)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            if k < 0:
                kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS))
            else:
                kk[xx, x] = int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _resample_axis0(img: np.ndarray, out_size: int) -> np.ndarray:
    """Resample along axis 0 of an [N, ...] uint8 array (one Pillow pass)."""
    bounds, kk, _ = precompute_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        k = kk[xx, :n].astype(np.int64).reshape((n,) + (1,) * (img.ndim - 1))
        ss = (1 << (PRECISION_BITS - 1)) + (src[x0:x0 + n] * k).sum(axis=0)
        out[xx] = np.clip(ss >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_resize_bicubic(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """``Image.fromarray(img).resize((out_w, out_h), BICUBIC)`` for an [H, W, C] uint8 array."""
    h, w = img.shape[:2]
    if (h, w) == (out_h, out_w):
        return img.copy()
    if w != out_w:  # horizontal pass first (ImagingResample), 8-bit intermediate
        img = np.ascontiguousarray(np.swapaxes(_resample_axis0(np.swapaxes(img, 0, 1), out_w), 0, 1))
    if h != out_h:
        img = _resample_axis0(img, out_h)
    return img


def resized_output_size(h: int, w: int, size: int):
    """torchvision ``_compute_resized_output_size`` for an int ``size`` (shorter edge → size)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)  # (new_h, new_w)


def crop_offsets(h: int, w: int, size: int):
    """torchvision ``center_crop`` offsets (valid when h, w >= size, which Resize(size) guarantees)."""
    return int(round((h - size) / 2.0)), int(round((w - size) / 2.0))


def resize_center_crop_u8(img: np.ndarray, size: int = 224) -> np.ndarray:
    """[H, W, 3] uint8 → [3, size, size] uint8: Resize(size, BICUBIC) + CenterCrop(size), channels first."""
    h, w = img.shape[:2]
    nh, nw = resized_output_size(h, w, size)
    r = pil_resize_bicubic(img, nh, nw)
    top, left = crop_offsets(nh, nw, size)
    return np.ascontiguousarray(r[top:top + size, left:left + size].transpose(2, 0, 1))


def to_tensor_normalize(u8_chw: np.ndarray, mean, std) -> np.ndarray:
    """ToTensor (÷255 in fp32) + Normalize ((x - mean) / std in fp32), eva_clip.py:141-152."""
    x = u8_chw.astype(np.float32) / np.float32(255.0)
    m = np.asarray(mean, np.float32).reshape(3, 1, 1)
    s = np.asarray(std, np.float32).reshape(3, 1, 1)
    return (x - m) / s
