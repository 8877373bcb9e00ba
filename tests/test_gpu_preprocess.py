"""Device preprocessing through the C ABI: bit-exact against golden vectors produced by
torchvision + Pillow, against the numpy oracle on random sizes, and through ``CLIPWrapper.get_preprocess()``."""
import glob
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR
from oracle.preprocess_oracle import preprocess as oracle_preprocess

pytestmark = pytest.mark.gpu
GOLDENS = sorted(glob.glob(os.path.join(GOLDEN_DIR, "preprocess_*.pt")))


@pytest.mark.parametrize("path", GOLDENS, ids=[os.path.basename(p)[:-3] for p in GOLDENS])
def test_bit_exact_vs_torchvision_golden(path):
    import tapclip_b200 as tb
    gold = torch.load(path, map_location="cpu", weights_only=False)
    out = tb.GpuPreprocess(gold["image_size"])(gold["image"])
    torch.cuda.synchronize()
    assert out.is_cuda and out.shape == gold["output"].shape
    assert torch.equal(out.cpu(), gold["output"])


def test_random_sizes_vs_oracle_and_cuda_input():
    import tapclip_b200 as tb
    g = np.random.default_rng(1)
    for i in range(12):
        h, w = int(g.integers(8, 400)), int(g.integers(8, 400))
        r = int(g.choice([16, 64, 224]))
        img = g.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        ref = torch.from_numpy(oracle_preprocess(img, r))
        pre = tb.GpuPreprocess(r)
        assert torch.equal(pre(img).cpu(), ref), (h, w, r)                               # numpy input
        assert torch.equal(pre(torch.from_numpy(img).cuda()).cpu(), ref), (h, w, r)      # CUDA uint8 tensor input


def test_pil_input_and_wrapper_surface():
    PIL_Image = pytest.importorskip("PIL.Image")
    import tapclip_b200 as tb
    clip = tb.CLIPWrapper("mini-16", None, "cuda", seed=0)
    pre = clip.get_preprocess()                                      # clip_wrapper.py:64-65
    assert pre is clip.preprocess and pre.image_size == 64
    img = np.random.default_rng(2).integers(0, 256, size=(90, 120, 3), dtype=np.uint8)
    out = pre(PIL_Image.fromarray(img))
    assert torch.equal(out.cpu(), torch.from_numpy(oracle_preprocess(img, 64)))
    gray = PIL_Image.fromarray(img[:, :, 0])                         # non-RGB modes are converted like open_clip's _convert_to_rgb
    assert pre(gray).shape == (3, 64, 64)
    feats = clip.encode_image(torch.stack([out, out]))               # the transform's output feeds encode_image directly
    assert feats.shape == (2, 256)
