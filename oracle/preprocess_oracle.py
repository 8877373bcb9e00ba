"""CPU restatement (numpy) of the image transform behind ``CLIPWrapper.get_preprocess()`` -- TEST INFRASTRUCTURE ONLY.

Reference call sites: models/clip_wrapper.py:13 (open_clip.create_model_and_transforms -> ``preprocess``), :64-65,
dataset.py:31 (``ImageFolder(transform=preprocess)``).  The arithmetic is third-party: open_clip's inference transform is
torchvision ``Resize(R, BICUBIC) -> CenterCrop(R) -> ToTensor -> Normalize(mean, std)`` and the resize is Pillow's
``ImagingResample`` (src/libImaging/Resample.c, 8 bits per channel).  Neither is pinned by the reference (no requirements
file); this image has Pillow 12.2.0 / torchvision 0.26.0, whose outputs are stored in tests/golden/preprocess_*.pt by
oracle/make_preprocess_goldens.py and compared bit for bit (tests/test_preprocess_oracle.py).
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
OPENAI_DATASET_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_DATASET_STD = (0.26862954, 0.26130258, 0.27577711)


def _bicubic(x: float) -> float:                      # Resample.c: bicubic_filter, a = -0.5
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c: precompute_coeffs + normalize_coeffs_8bpc (box = whole axis).  Returns (bounds [out,2], kk [out,ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bicubic(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Pillow Image.resize((out_w, out_h), BICUBIC) for an [H, W, 3] uint8 image: horizontal pass, then vertical pass."""
    h, w, _ = img.shape
    src = img.astype(np.int64)
    if out_w != w:
        bounds, kk = precompute_coeffs(w, out_w)
        tmp = np.empty((h, out_w, 3), dtype=np.uint8)
        for xx in range(out_w):
            x0, n = bounds[xx]
            acc = (src[:, x0:x0 + n, :] * kk[xx, :n, None]).sum(axis=1) + (1 << (PRECISION_BITS - 1))
            tmp[:, xx, :] = _clip8(acc)
        src = tmp.astype(np.int64)
    else:
        tmp = img
    if out_h != h:
        bounds, kk = precompute_coeffs(h, out_h)
        out = np.empty((out_h, src.shape[1], 3), dtype=np.uint8)
        for yy in range(out_h):
            y0, n = bounds[yy]
            acc = (src[y0:y0 + n, :, :] * kk[yy, :n, None, None]).sum(axis=0) + (1 << (PRECISION_BITS - 1))
            out[yy] = _clip8(acc)
        return out
    return np.ascontiguousarray(tmp)


def preprocess(img: np.ndarray, image_size: int, mean=OPENAI_DATASET_MEAN, std=OPENAI_DATASET_STD) -> np.ndarray:
    """[H, W, 3] uint8 RGB -> [3, R, R] float32, as torchvision's Resize(R, BICUBIC) / CenterCrop / ToTensor / Normalize."""
    h, w, _ = img.shape
    r = image_size
    oh, ow = (int(r * h / w), r) if w <= h else (r, int(r * w / h))            # _compute_resized_output_size
    res = resize_bicubic(img, oh, ow)
    top, left = int(round((oh - r) / 2.0)), int(round((ow - r) / 2.0))           # center_crop
    crop = res[top:top + r, left:left + r, :]
    x = crop.astype(np.float32).transpose(2, 0, 1) / np.float32(255)             # ToTensor
    m = np.asarray(mean, dtype=np.float32)[:, None, None]
    s = np.asarray(std, dtype=np.float32)[:, None, None]
    return ((x - m) / s).astype(np.float32)                                      # Normalize
