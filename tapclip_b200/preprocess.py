"""Device-side replacement for the transform behind ``CLIPWrapper.get_preprocess()`` (models/clip_wrapper.py:13,64-65;
applied per image at dataset.py:31): open_clip's inference transform

    Resize(R, BICUBIC) on the shorter side -> CenterCrop(R) -> ToTensor -> Normalize(OPENAI_MEAN, OPENAI_STD)

run by ``tapclip_op_preprocess`` (csrc/preprocess.cu), bit-exact with Pillow's 8-bit antialiased bicubic resampling.
Input: a ``PIL.Image`` (any mode, converted to RGB like open_clip's ``_convert_to_rgb``), or a uint8 ``[H, W, 3]`` RGB
numpy array / torch tensor (CPU or CUDA).  Output: fp32 CUDA tensor ``[3, R, R]``.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

OPENAI_DATASET_MEAN = (0.48145466, 0.4578275, 0.40821073)        # open_clip constants
OPENAI_DATASET_STD = (0.26862954, 0.26130258, 0.27577711)


class GpuPreprocess:
    def __init__(self, image_size: int, mean=OPENAI_DATASET_MEAN, std=OPENAI_DATASET_STD, device="cuda"):
        self.image_size = int(image_size)
        self.device = torch.device(device)
        self._mean = (C.c_float * 3)(*mean)
        self._std = (C.c_float * 3)(*std)

    def resized_size(self, h: int, w: int):
        """torchvision ``_compute_resized_output_size`` for ``size=[R]``: (out_h, out_w)."""
        r = self.image_size
        return (int(r * h / w), r) if w <= h else (r, int(r * w / h))

    def __call__(self, img) -> torch.Tensor:
        if not isinstance(img, torch.Tensor):
            if hasattr(img, "convert"):                         # PIL.Image
                img = img.convert("RGB")
            import numpy as np
            img = torch.from_numpy(np.array(img, copy=True))       # writable copy (PIL exposes read-only buffers)
        if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3:
            raise ValueError(f"expected a uint8 [H, W, 3] RGB image, got {tuple(img.shape)} {img.dtype}")
        if self.device.type != "cuda":
            raise _lib.TapclipError("GpuPreprocess needs a CUDA device (tapclip_b200 has no CPU fallback)")
        img = img.to(self.device, non_blocking=True).contiguous()
        h, w = int(img.shape[0]), int(img.shape[1])
        r = self.image_size
        oh, ow = self.resized_size(h, w)
        top, left = int(round((oh - r) / 2.0)), int(round((ow - r) / 2.0))     # torchvision center_crop (round half to even)
        out = torch.empty(3, r, r, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().tapclip_op_preprocess(_lib.ptr(img), h, w, _lib.ptr(out), r, top, left,
                                                         C.cast(self._mean, C.c_void_p), C.cast(self._std, C.c_void_p), _lib.stream_ptr()))
        return out

    def __repr__(self):
        return f"GpuPreprocess(image_size={self.image_size})"
