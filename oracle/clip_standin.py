"""Open_clip-shaped CLIP stand-in + the reference's CLIPWrapper, restated (TEST INFRASTRUCTURE).

The reference builds its model with ``open_clip.create_model_and_transforms``
(``models/clip_wrapper.py:13``); open_clip is an un-vendored, un-pinned
dependency that is absent from this image, so its model is restated here from
its published structure with stock ``torch.nn`` modules:

    CLIP{visual, transformer, token_embedding, positional_embedding, ln_final,
         text_projection, logit_scale}
    VisionTransformer{conv1(bias=False), class_embedding, positional_embedding,
         ln_pre, transformer, ln_post, proj}
    Transformer.resblocks[i] = {ln_1, attn=nn.MultiheadAttention(batch_first),
         ln_2, mlp=Sequential(c_fc, gelu, c_proj)}
    x = x + attn(ln_1(x), need_weights=False)[0];  x = x + mlp(ln_2(x))

State-dict key names equal open_clip's, so a real open_clip checkpoint loads.

``StandInCLIPWrapper`` follows ``models/clip_wrapper.py:9-65`` line by line
except for the two open_clip calls (model construction, tokenizer).

Attribution modes:

* ``literal``  – the hooked ``attn`` module returns ``(attn_output, None)`` as
  stock open_clip does (``need_weights=False``); the reference hook then
  stores a token-mean of features and the attribution degenerates to 1.0.
* ``intended`` – the hooked module returns the per-head attention
  probabilities ``[B,H,T,T]`` first, which is what the hook's own comments
  (``clip_wrapper.py:35-36``) say it receives.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from dataclasses import dataclass, replace

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# configurations (open_clip model_configs/*.json values)
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class CLIPConfig:
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


_BASE = {
    "ViT-B-32": CLIPConfig("ViT-B-32", 512, 224, 32, 768, 12, 12, 512, 12, 8),
    "ViT-B-16": CLIPConfig("ViT-B-16", 512, 224, 16, 768, 12, 12, 512, 12, 8),
    "ViT-L-14": CLIPConfig("ViT-L-14", 768, 224, 14, 1024, 24, 16, 768, 12, 12),
    "ViT-L-14-336": CLIPConfig("ViT-L-14-336", 768, 336, 14, 1024, 24, 16, 768, 12, 12),
    # small shapes for tests (same code path, head_dim 64 everywhere)
    "mini-16": CLIPConfig("mini-16", 256, 64, 16, 256, 2, 4, 256, 2, 4),
    "mini-14": CLIPConfig("mini-14", 128, 56, 14, 128, 3, 2, 128, 2, 2),
    "mini-n197": CLIPConfig("mini-n197", 128, 224, 16, 128, 2, 2, 128, 2, 2),   # ViT-B/16 token count (197) on a tiny width
    "mini-t512": CLIPConfig("mini-t512", 256, 64, 16, 256, 2, 4, 512, 2, 8),   # text width 512: PromptAdjustor 'residual' hard-codes it
}


def get_config(model_name: str) -> CLIPConfig:
    """open_clip naming: a ``-quickgelu`` suffix selects x*sigmoid(1.702x), plain names nn.GELU (erf)."""
    quick = model_name.endswith("-quickgelu")
    base = model_name[: -len("-quickgelu")] if quick else model_name
    if base not in _BASE:
        raise ValueError(f"unknown model_name {model_name!r}; known: {sorted(_BASE)}")
    return replace(_BASE[base], name=model_name, quick_gelu=quick)


# --------------------------------------------------------------------------------------
# model
# --------------------------------------------------------------------------------------
class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class HookableMHA(nn.MultiheadAttention):
    """nn.MultiheadAttention whose module OUTPUT (what a forward hook sees) is selectable.

    emit == "output":  (attn_output, None)                       — stock open_clip behaviour
    emit == "weights": (per-head probabilities [B,H,T,T], attn_output)
    """

    emit = "output"

    def forward(self, x, attn_mask=None):  # type: ignore[override]
        if self.emit == "output":
            return super().forward(x, x, x, need_weights=False, attn_mask=attn_mask)
        out, w = super().forward(x, x, x, need_weights=True, average_attn_weights=False, attn_mask=attn_mask)
        return w, out


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model: int, n_head: int, quick_gelu: bool):
        super().__init__()
        self.ln_1 = nn.LayerNorm(d_model)
        self.attn = HookableMHA(d_model, n_head, batch_first=True)
        self.ln_2 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", QuickGELU() if quick_gelu else nn.GELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model)),
        ]))

    def forward(self, x, attn_mask=None):
        y = self.attn(self.ln_1(x), attn_mask=attn_mask)
        x = x + (y[0] if self.attn.emit == "output" else y[1])
        x = x + self.mlp(self.ln_2(x))
        return x


class Transformer(nn.Module):
    """batch-first [S, T, D] in and out (open_clip >= 2.24 ``batch_first=True``)."""

    def __init__(self, width: int, layers: int, heads: int, quick_gelu: bool):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.ModuleList([ResidualAttentionBlock(width, heads, quick_gelu) for _ in range(layers)])

    def forward(self, x, attn_mask=None):
        for blk in self.resblocks:
            x = blk(x, attn_mask=attn_mask)
        return x


class VisionTransformer(nn.Module):
    def __init__(self, cfg: CLIPConfig):
        super().__init__()
        d = cfg.vision_width
        self.conv1 = nn.Conv2d(3, d, kernel_size=cfg.patch_size, stride=cfg.patch_size, bias=False)
        scale = d ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(d))
        self.positional_embedding = nn.Parameter(scale * torch.randn(cfg.vision_tokens, d))
        self.ln_pre = nn.LayerNorm(d)
        self.transformer = Transformer(d, cfg.vision_layers, cfg.vision_heads, cfg.quick_gelu)
        self.ln_post = nn.LayerNorm(d)
        self.proj = nn.Parameter(scale * torch.randn(d, cfg.embed_dim))

    def forward(self, x):
        x = self.conv1(x)                                   # [B, d, g, g]
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)   # [B, g*g, d]
        cls = self.class_embedding.to(x.dtype).expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], dim=1) + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        x = self.transformer(x)
        pooled = self.ln_post(x[:, 0])
        return pooled @ self.proj


class CLIP(nn.Module):
    def __init__(self, cfg: CLIPConfig):
        super().__init__()
        self.cfg = cfg
        self.context_length = cfg.context_length
        self.visual = VisionTransformer(cfg)
        self.transformer = Transformer(cfg.text_width, cfg.text_layers, cfg.text_heads, cfg.quick_gelu)
        self.vocab_size = cfg.vocab_size
        self.token_embedding = nn.Embedding(cfg.vocab_size, cfg.text_width)
        self.positional_embedding = nn.Parameter(torch.empty(cfg.context_length, cfg.text_width))
        self.ln_final = nn.LayerNorm(cfg.text_width)
        self.text_projection = nn.Parameter(torch.empty(cfg.text_width, cfg.embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))
        mask = torch.full((cfg.context_length, cfg.context_length), float("-inf")).triu_(1)
        self.register_buffer("attn_mask", mask, persistent=False)

    def encode_image(self, image):
        return self.visual(image)

    def encode_text(self, text):
        """Standard CLIP text path (NOT what FullModel uses)."""
        x = self.token_embedding(text) + self.positional_embedding
        x = self.transformer(x, attn_mask=self.attn_mask)
        x = self.ln_final(x)
        return x[torch.arange(x.shape[0]), text.argmax(dim=-1)] @ self.text_projection


def init_clip_weights(model: CLIP, seed: int = 0) -> CLIP:
    """Seeded open_clip-style random init.  One generator, fixed parameter order."""
    g = torch.Generator().manual_seed(seed)

    def normal_(t, std):
        with torch.no_grad():
            t.copy_(torch.randn(t.shape, generator=g) * std)

    def uniform_(t, bound):
        with torch.no_grad():
            t.copy_((torch.rand(t.shape, generator=g) * 2 - 1) * bound)

    cfg = model.cfg
    normal_(model.token_embedding.weight, 0.02)
    normal_(model.positional_embedding, 0.01)
    D, L = cfg.text_width, cfg.text_layers
    proj_std, attn_std, fc_std = (D ** -0.5) * ((2 * L) ** -0.5), D ** -0.5, (2 * D) ** -0.5
    for blk in model.transformer.resblocks:
        normal_(blk.attn.in_proj_weight, attn_std)
        normal_(blk.attn.in_proj_bias, 0.02)
        normal_(blk.attn.out_proj.weight, proj_std)
        normal_(blk.attn.out_proj.bias, 0.02)
        normal_(blk.mlp.c_fc.weight, fc_std)
        normal_(blk.mlp.c_fc.bias, 0.02)
        normal_(blk.mlp.c_proj.weight, proj_std)
        normal_(blk.mlp.c_proj.bias, 0.02)
        for ln in (blk.ln_1, blk.ln_2):
            normal_(ln.weight, 0.05); ln.weight.data += 1.0
            normal_(ln.bias, 0.05)
    normal_(model.ln_final.weight, 0.05); model.ln_final.weight.data += 1.0
    normal_(model.ln_final.bias, 0.05)
    normal_(model.text_projection, D ** -0.5)

    v = model.visual
    d = cfg.vision_width
    fan_in = 3 * cfg.patch_size ** 2
    uniform_(v.conv1.weight, fan_in ** -0.5)
    normal_(v.class_embedding, d ** -0.5)
    normal_(v.positional_embedding, d ** -0.5)
    normal_(v.proj, d ** -0.5)
    for blk in v.transformer.resblocks:
        uniform_(blk.attn.in_proj_weight, (6.0 / (4 * d)) ** 0.5)       # xavier_uniform of [3d, d]
        normal_(blk.attn.in_proj_bias, 0.02)
        uniform_(blk.attn.out_proj.weight, d ** -0.5)
        normal_(blk.attn.out_proj.bias, 0.02)
        uniform_(blk.mlp.c_fc.weight, d ** -0.5)
        uniform_(blk.mlp.c_fc.bias, d ** -0.5)
        uniform_(blk.mlp.c_proj.weight, (4 * d) ** -0.5)
        uniform_(blk.mlp.c_proj.bias, (4 * d) ** -0.5)
        for ln in (blk.ln_1, blk.ln_2):
            normal_(ln.weight, 0.05); ln.weight.data += 1.0
            normal_(ln.bias, 0.05)
    for ln in (v.ln_pre, v.ln_post):
        normal_(ln.weight, 0.05); ln.weight.data += 1.0
        normal_(ln.bias, 0.05)
    return model


def build_clip(model_name: str, seed: int = 0) -> CLIP:
    model = CLIP(get_config(model_name))
    return init_clip_weights(model, seed).eval()


# --------------------------------------------------------------------------------------
# tokenizer (the BPE vocabulary ships inside open_clip and is unavailable)
# --------------------------------------------------------------------------------------
class SyntheticTokenizer:
    """``tokenizer(str) -> LongTensor[1, 77]`` with SOT 49406, EOT 49407, pad 0.

    "a photo of a <class>" becomes [SOT, 320, 1125, 539, 320, <1-3 ids derived from a CRC of
    the class words>, EOT, 0...].  Deterministic across processes (no Python hash()).
    """

    SOT, EOT = 49406, 49407
    WORDS = {"a": 320, "photo": 1125, "of": 539}

    def __init__(self, context_length: int = 77):
        self.context_length = context_length

    def _word_ids(self, word: str):
        if word in self.WORDS:
            return [self.WORDS[word]]
        h = zlib.crc32(word.encode("utf-8"))
        n = 1 + h % 3
        return [1000 + (zlib.crc32(f"{word}#{i}".encode("utf-8")) % 48405) for i in range(n)]

    def __call__(self, texts, context_length: int | None = None):
        if isinstance(texts, str):
            texts = [texts]
        L = context_length or self.context_length
        out = torch.zeros(len(texts), L, dtype=torch.long)
        for i, t in enumerate(texts):
            ids = [self.SOT]
            for w in t.lower().split():
                ids.extend(self._word_ids(w))
            ids = ids[: L - 1] + [self.EOT]
            out[i, : len(ids)] = torch.tensor(ids)
        return out


# --------------------------------------------------------------------------------------
# CLIPWrapper restated (models/clip_wrapper.py:9-65)
# --------------------------------------------------------------------------------------
class StandInCLIPWrapper(nn.Module):
    """Same attributes/methods as the reference ``CLIPWrapper``; ``model`` is the stand-in CLIP.

    ``attribution`` selects what the hooked module emits (see module docstring).
    ``state_dict`` (instead of ``pretrained_path``) seeds the weights; both are accepted.
    """

    def __init__(self, model_name="ViT-B-32", pretrained_path=None, device="cpu", *, seed=0,
                 state_dict=None, attribution="literal"):
        super().__init__()
        if attribution not in ("literal", "intended"):
            raise ValueError(f"Unknown attribution mode: {attribution}")
        self.device = device
        self.attribution = attribution
        self.model = CLIP(get_config(model_name))          # clip_wrapper.py:13 (pretrained='')
        self.preprocess = None                              # torchvision transform in the reference; host I/O, out of scope
        if pretrained_path is not None:                     # clip_wrapper.py:14-15
            state_dict = torch.load(pretrained_path, map_location=device)
        if state_dict is not None:
            self.model.load_state_dict(state_dict, strict=True)
        else:
            init_clip_weights(self.model, seed)
        self.model.to(device).eval()                        # clip_wrapper.py:16
        for p in self.model.parameters():                   # clip_wrapper.py:19-20
            p.requires_grad = False
        self.attention_maps = []                            # clip_wrapper.py:23
        self._register_text_attention_hook()                # clip_wrapper.py:24
        self.tokenizer = SyntheticTokenizer(self.model.context_length)   # clip_wrapper.py:27

    def _register_text_attention_hook(self):                # clip_wrapper.py:29-40
        def hook_fn(module, input, output):
            attn = output[0].detach().mean(dim=1)
            self.attention_maps.append(attn)

        last_text_block = self.model.transformer.resblocks[-1].attn
        if self.attribution == "intended":
            last_text_block.emit = "weights"
        last_text_block.register_forward_hook(hook_fn)

    def reset(self):                                        # clip_wrapper.py:42-44
        self.attention_maps.clear()

    def encode_image(self, image_tensor):                   # clip_wrapper.py:46-47
        return self.model.encode_image(image_tensor)

    def encode_text(self, token_tensor):                    # clip_wrapper.py:49-51
        self.reset()
        return self.model.encode_text(token_tensor)

    def get_attention_map(self):                            # clip_wrapper.py:53-59
        if len(self.attention_maps) == 0:
            return None
        return self.attention_maps[-1]

    def get_tokenizer(self):                                # clip_wrapper.py:61-62
        return self.tokenizer

    def get_preprocess(self):                               # clip_wrapper.py:64-65
        return self.preprocess


# --------------------------------------------------------------------------------------
# image-side attention attribution (north-star extension; NOT in the reference; parity unpinned)
# --------------------------------------------------------------------------------------
@torch.no_grad()
def vision_cls_attention(model: CLIP, images: torch.Tensor):
    """Per-layer, per-head CLS-row attention probabilities of the vision tower.

    Uses the same forward-hook mechanism as clip_wrapper.py:29-40, on every
    ``visual.transformer.resblocks[l].attn``.  Returns (features [B,E], cls_rows [B,L,H,N]).
    """
    blocks = model.visual.transformer.resblocks
    rows, handles, saved = [], [], []
    for blk in blocks:
        saved.append(blk.attn.emit)
        blk.attn.emit = "weights"
        handles.append(blk.attn.register_forward_hook(lambda m, i, o: rows.append(o[0][:, :, 0, :].detach())))
    try:
        feats = model.encode_image(images)
    finally:
        for blk, h, e in zip(blocks, handles, saved):
            h.remove()
            blk.attn.emit = e
    return feats, torch.stack(rows, dim=1)


@torch.no_grad()
def vision_attention_rollout(model: CLIP, images: torch.Tensor):
    """Attention rollout (Abnar & Zuidema 2020): R = prod_l rownorm(0.5*mean_h P_l + 0.5*I); returns R[:,0,1:]."""
    blocks = model.visual.transformer.resblocks
    maps, handles, saved = [], [], []
    for blk in blocks:
        saved.append(blk.attn.emit)
        blk.attn.emit = "weights"
        handles.append(blk.attn.register_forward_hook(lambda m, i, o: maps.append(o[0].detach().mean(dim=1))))
    try:
        model.encode_image(images)
    finally:
        for blk, h, e in zip(blocks, handles, saved):
            h.remove()
            blk.attn.emit = e
    n = maps[0].shape[-1]
    eye = torch.eye(n)
    R = eye.expand(images.shape[0], n, n).clone()
    for A in maps:
        A = 0.5 * A + 0.5 * eye
        A = A / A.sum(dim=-1, keepdim=True)
        R = A @ R
    return R[:, 0, 1:]
