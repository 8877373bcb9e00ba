"""Generate tests/golden/preprocess_*.pt with the real libraries behind the reference's preprocess (torchvision + Pillow,
as installed in this image) and check the numpy restatement against them bit for bit.

Run:  python -m oracle.make_preprocess_goldens
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.preprocess_oracle import OPENAI_DATASET_MEAN, OPENAI_DATASET_STD, preprocess     # noqa: E402

# name, H, W, R: landscape / portrait down-scaling, identity, up-scaling, odd crop offsets (round-half-even), full size
CASES = [("land", 150, 200, 64), ("port", 211, 140, 64), ("ident", 64, 64, 64), ("up", 40, 57, 64), ("odd", 67, 64, 64),
         ("b16", 240, 331, 224)]


def synthetic_image(h, w, seed):
    """Smooth gradients + texture + noise (exercises negative bicubic lobes and clipping)."""
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    base = np.stack([127 + 120 * np.sin(xx / 7.0 + seed) * np.cos(yy / 11.0), 255 * (xx / max(w - 1, 1)), 255 * ((xx + yy) % 17 < 8)], -1)
    return np.clip(base + g.normal(0, 25, size=(h, w, 3)), 0, 255).astype(np.uint8)


def torchvision_reference(img, r):
    from PIL import Image
    from torchvision import transforms as T
    from torchvision.transforms import InterpolationMode
    tf = T.Compose([T.Resize(r, interpolation=InterpolationMode.BICUBIC), T.CenterCrop(r), T.ToTensor(),
                    T.Normalize(OPENAI_DATASET_MEAN, OPENAI_DATASET_STD)])          # open_clip image_transform(is_train=False)
    return tf(Image.fromarray(img))


def main():
    import PIL
    import torchvision
    out_dir = os.path.join(ROOT, "tests", "golden")
    for i, (name, h, w, r) in enumerate(CASES):
        img = synthetic_image(h, w, i)
        ref = torchvision_reference(img, r)
        mine = torch.from_numpy(preprocess(img, r))
        assert torch.equal(mine, ref), f"{name}: numpy restatement differs from torchvision/Pillow, max {(mine - ref).abs().max()}"
        torch.save({"case": name, "image": torch.from_numpy(img), "image_size": r, "output": ref.clone(),
                    "pillow": PIL.__version__, "torchvision": torchvision.__version__},
                   os.path.join(out_dir, f"preprocess_{name}.pt"))
        print(f"preprocess_{name}: {h}x{w} -> {r}: restatement bit-exact")


if __name__ == "__main__":
    main()
