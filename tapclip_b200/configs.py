"""Model shapes the engine is built for (values of open_clip's model_configs/*.json).

open_clip selects the activation by model name: a ``-quickgelu`` suffix means x*sigmoid(1.702x),
plain names mean nn.GELU (erf) — the reference passes the name straight through
(models/clip_wrapper.py:13), so the same rule applies here.
"""
from __future__ import annotations

from dataclasses import dataclass, replace


@dataclass(frozen=True)
class ModelConfig:
    name: str
    embed_dim: int
    image_size: int
    patch_size: int
    vision_width: int
    vision_layers: int
    vision_heads: int
    text_width: int
    text_layers: int
    text_heads: int
    context_length: int = 77
    vocab_size: int = 49408
    quick_gelu: bool = False

    @property
    def vision_tokens(self) -> int:
        return (self.image_size // self.patch_size) ** 2 + 1


_TABLE = {
    "ViT-B-32": ModelConfig("ViT-B-32", 512, 224, 32, 768, 12, 12, 512, 12, 8),
    "ViT-B-16": ModelConfig("ViT-B-16", 512, 224, 16, 768, 12, 12, 512, 12, 8),
    "ViT-L-14": ModelConfig("ViT-L-14", 768, 224, 14, 1024, 24, 16, 768, 12, 12),
    "ViT-L-14-336": ModelConfig("ViT-L-14-336", 768, 336, 14, 1024, 24, 16, 768, 12, 12),
    "mini-16": ModelConfig("mini-16", 256, 64, 16, 256, 2, 4, 256, 2, 4),
    "mini-14": ModelConfig("mini-14", 128, 56, 14, 128, 3, 2, 128, 2, 2),
    "mini-n197": ModelConfig("mini-n197", 128, 224, 16, 128, 2, 2, 128, 2, 2),   # ViT-B/16 token count (197) on a tiny width
    "mini-t512": ModelConfig("mini-t512", 256, 64, 16, 256, 2, 4, 512, 2, 8),   # text width 512 (PromptAdjustor 'residual')
}


def get_model_config(model_name: str) -> ModelConfig:
    quick = model_name.endswith("-quickgelu")
    base = model_name[: -len("-quickgelu")] if quick else model_name
    if base not in _TABLE:
        raise ValueError(f"Unknown model_name: {model_name!r} (known: {sorted(_TABLE)})")
    return replace(_TABLE[base], name=model_name, quick_gelu=quick)


def flops_per_image(cfg: ModelConfig) -> float:
    """Algorithmic forward FLOPs (2*MAC) of the vision tower for one image."""
    n, d = cfg.vision_tokens, cfg.vision_width
    per_layer = 24 * n * d * d + 4 * n * n * d
    patch = 2 * (n - 1) * (3 * cfg.patch_size ** 2) * d
    return cfg.vision_layers * per_layer + patch + 2 * d * cfg.embed_dim


def flops_per_text_sequence(cfg: ModelConfig, seq_len: int) -> float:
    """Algorithmic forward FLOPs of the text transformer for one [T, D] sequence (+ projection)."""
    t, d = seq_len, cfg.text_width
    return cfg.text_layers * (24 * t * d * d + 4 * t * t * d) + 2 * d * cfg.embed_dim
