"""The numpy restatement of the reference's preprocess (open_clip inference transform = torchvision Resize-BICUBIC /
CenterCrop / ToTensor / Normalize over Pillow's resampler) against golden vectors produced by the real libraries
(oracle/make_preprocess_goldens.py), and against the libraries themselves where they are installed."""
import glob
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR
from oracle.preprocess_oracle import precompute_coeffs, preprocess

GOLDENS = sorted(glob.glob(os.path.join(GOLDEN_DIR, "preprocess_*.pt")))


def test_golden_files_exist():
    assert len(GOLDENS) >= 6


@pytest.mark.parametrize("path", GOLDENS, ids=[os.path.basename(p)[:-3] for p in GOLDENS])
def test_restatement_is_bit_exact_vs_torchvision_golden(path):
    gold = torch.load(path, map_location="cpu", weights_only=False)
    out = torch.from_numpy(preprocess(gold["image"].numpy(), gold["image_size"]))
    assert out.shape == (3, gold["image_size"], gold["image_size"]) and out.dtype == torch.float32
    assert torch.equal(out, gold["output"])


def test_coefficients_are_normalised_fixed_point():
    for n_in, n_out in ((200, 64), (64, 64), (57, 64), (1000, 224)):
        bounds, kk = precompute_coeffs(n_in, n_out)
        assert (bounds[:, 0] >= 0).all() and (bounds[:, 0] + bounds[:, 1] <= n_in).all()
        assert np.abs(kk.sum(axis=1) - (1 << 22)).max() <= kk.shape[1]          # weights sum to 1.0 up to per-tap rounding


def test_restatement_matches_live_torchvision_on_random_sizes():
    tv = pytest.importorskip("torchvision")
    pytest.importorskip("PIL")
    from oracle.make_preprocess_goldens import synthetic_image, torchvision_reference
    g = np.random.default_rng(0)
    for i in range(6):
        h, w = int(g.integers(20, 260)), int(g.integers(20, 260))
        r = int(g.choice([32, 56, 64]))
        img = synthetic_image(h, w, 100 + i)
        assert torch.equal(torch.from_numpy(preprocess(img, r)), torchvision_reference(img, r)), (h, w, r)


def test_gpu_preprocess_fails_loudly_without_cuda():
    import tapclip_b200 as tb
    from tapclip_b200._lib import TapclipError
    p = tb.GpuPreprocess(64, device="cpu")
    assert p.resized_size(150, 200) == (64, 85) and p.resized_size(211, 140) == (96, 64)
    with pytest.raises(TapclipError):
        p(torch.zeros(10, 10, 3, dtype=torch.uint8))
    with pytest.raises(ValueError):
        p(torch.zeros(10, 10, 4, dtype=torch.uint8))
