"""Drop-in for the reference's ``models/clip_wrapper.py`` (CLIPWrapper, lines 9-65) on B200.

Same constructor signature, attributes and methods; what changes is what sits underneath:
``open_clip.create_model_and_transforms`` + PyTorch eager kernels are replaced by the libtapclip
engine (hand-written sm_100a kernels behind the C ABI of include/tapclip.h).

* ``.model`` keeps open_clip's attribute / state-dict names (``visual.*``, ``transformer.resblocks.*``,
  ``token_embedding``, ``text_projection`` ...) as frozen fp32 parameters, so checkpoints written by
  train.py:131-132 and read by test_cross_domain.py:43-61 stay interchangeable.  The engine holds
  its own converted (bf16 / transposed) copies; they are refreshed after every ``load_state_dict``.
* The forward hook of clip_wrapper.py:29-40 does not exist: the probabilities it was meant to capture
  are emitted by the attention kernel's probe epilogue.  ``attention_maps`` therefore holds the compact
  per-class probe (``[C, P]`` head-mean of column T-1 for the P ctx rows, exactly the slice
  attribution_monitor.py:29 takes) instead of full ``[B,T,T]`` maps; see DESIGN.md.
* ``attribution`` ('literal' | 'intended') states what the hooked module would have emitted
  (SURVEY.md fact 6): 'literal' = stock open_clip (``need_weights=False`` -> attribution == 1.0),
  'intended' = per-head attention probabilities, as the hook's comments say.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .configs import ModelConfig, get_model_config
from .engine import Engine
from .preprocess import GpuPreprocess
from .tokenizer import SyntheticTokenizer


class _Params(nn.Module):
    """Attribute container that only carries parameters (open_clip module names, no torch compute)."""


def _ln(d):
    m = _Params()
    m.weight = nn.Parameter(torch.ones(d))
    m.bias = nn.Parameter(torch.zeros(d))
    return m


def _linear(n, k):
    m = _Params()
    m.weight = nn.Parameter(torch.empty(n, k))
    m.bias = nn.Parameter(torch.zeros(n))
    return m


def _resblock(d):
    b = _Params()
    b.ln_1 = _ln(d)
    b.attn = _Params()
    b.attn.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
    b.attn.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
    b.attn.out_proj = _linear(d, d)
    b.ln_2 = _ln(d)
    b.mlp = _Params()
    b.mlp.c_fc = _linear(4 * d, d)
    b.mlp.c_proj = _linear(d, 4 * d)
    return b


class _Tower(_Params):
    def __init__(self, d, layers):
        super().__init__()
        self.resblocks = nn.ModuleList([_resblock(d) for _ in range(layers)])


class EngineCLIP(nn.Module):
    """Parameter skeleton with open_clip's names; compute goes through the engine."""

    def __init__(self, cfg: ModelConfig):
        super().__init__()
        self.cfg = cfg
        self.context_length, self.vocab_size = cfg.context_length, cfg.vocab_size
        d = cfg.vision_width
        v = _Params()
        v.conv1 = _Params()
        v.conv1.weight = nn.Parameter(torch.empty(d, 3, cfg.patch_size, cfg.patch_size))
        v.class_embedding = nn.Parameter(torch.empty(d))
        v.positional_embedding = nn.Parameter(torch.empty(cfg.vision_tokens, d))
        v.ln_pre, v.ln_post = _ln(d), _ln(d)
        v.transformer = _Tower(d, cfg.vision_layers)
        v.proj = nn.Parameter(torch.empty(d, cfg.embed_dim))
        self.visual = v
        self.transformer = _Tower(cfg.text_width, cfg.text_layers)
        self.token_embedding = nn.Embedding(cfg.vocab_size, cfg.text_width)
        self.positional_embedding = nn.Parameter(torch.empty(cfg.context_length, cfg.text_width))
        self.ln_final = _ln(cfg.text_width)
        self.text_projection = nn.Parameter(torch.empty(cfg.text_width, cfg.embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))
        self._engine = None

    def encode_image(self, image):
        """[B,3,R,R] -> [B,E] un-normalised features (open_clip CLIP.encode_image, normalize=False)."""
        return self._engine.encode_image(image.contiguous().float())

    def image_attribution(self, image, rollout=True):
        """North-star extension (not in the reference): image features plus CLS attention attribution.
        Returns (features [B,E], cls_rows [B,L,H,N], rollout [B,N-1] or None)."""
        out = self._engine.encode_image(image.contiguous().float(), want_cls_rows=True, want_rollout=rollout)
        return out if rollout else (out[0], out[1], None)

    def encode_text(self, text):
        """Standard CLIP text path (open_clip CLIP.encode_text, normalize=False): int64 token ids [S, 77] -> [S, E]."""
        return self._engine.encode_text(text.to(self.text_projection.device))


@torch.no_grad()
def random_init_(model: EngineCLIP, seed: int = 0):
    """open_clip-style random init directly on the parameters' device (synthetic weights for benchmarks)."""
    cfg = model.cfg
    g = torch.Generator(device="cpu").manual_seed(seed)

    def normal_(p, std, mean=0.0):
        p.copy_((torch.randn(p.shape, generator=g) * std + mean).to(p.device))

    def uniform_(p, bound):
        p.copy_(((torch.rand(p.shape, generator=g) * 2 - 1) * bound).to(p.device))

    D, L = cfg.text_width, cfg.text_layers
    normal_(model.token_embedding.weight, 0.02)
    normal_(model.positional_embedding, 0.01)
    normal_(model.text_projection, D ** -0.5)
    normal_(model.ln_final.weight, 0.05, 1.0)
    normal_(model.ln_final.bias, 0.05)
    for blk in model.transformer.resblocks:
        normal_(blk.attn.in_proj_weight, D ** -0.5); normal_(blk.attn.in_proj_bias, 0.02)
        normal_(blk.attn.out_proj.weight, (D ** -0.5) * ((2 * L) ** -0.5)); normal_(blk.attn.out_proj.bias, 0.02)
        normal_(blk.mlp.c_fc.weight, (2 * D) ** -0.5); normal_(blk.mlp.c_fc.bias, 0.02)
        normal_(blk.mlp.c_proj.weight, (D ** -0.5) * ((2 * L) ** -0.5)); normal_(blk.mlp.c_proj.bias, 0.02)
        for ln in (blk.ln_1, blk.ln_2):
            normal_(ln.weight, 0.05, 1.0); normal_(ln.bias, 0.05)
    v, d = model.visual, cfg.vision_width
    uniform_(v.conv1.weight, (3 * cfg.patch_size ** 2) ** -0.5)
    normal_(v.class_embedding, d ** -0.5); normal_(v.positional_embedding, d ** -0.5); normal_(v.proj, d ** -0.5)
    for ln in (v.ln_pre, v.ln_post):
        normal_(ln.weight, 0.05, 1.0); normal_(ln.bias, 0.05)
    for blk in v.transformer.resblocks:
        uniform_(blk.attn.in_proj_weight, (6.0 / (4 * d)) ** 0.5); normal_(blk.attn.in_proj_bias, 0.02)
        uniform_(blk.attn.out_proj.weight, d ** -0.5); normal_(blk.attn.out_proj.bias, 0.02)
        uniform_(blk.mlp.c_fc.weight, d ** -0.5); uniform_(blk.mlp.c_fc.bias, d ** -0.5)
        uniform_(blk.mlp.c_proj.weight, (4 * d) ** -0.5); uniform_(blk.mlp.c_proj.bias, (4 * d) ** -0.5)
        for ln in (blk.ln_1, blk.ln_2):
            normal_(ln.weight, 0.05, 1.0); normal_(ln.bias, 0.05)


class CLIPWrapper(nn.Module):
    """models/clip_wrapper.py:9-65, same surface.

    Extra keyword-only arguments (reference-compatible defaults):
      state_dict   open_clip-named weights instead of ``pretrained_path``
      seed         random-init seed when neither a path nor a state_dict is given (synthetic weights)
      attribution  'literal' | 'intended' (see module docstring)
      dtype        'mixed' (default: tcgen05 tensor cores, bf16 operands in the image tower and the backward pass,
                   fp16 operands in the text-tower forward — meets the 1e-2 logit bar) | 'bf16' (bf16 operands
                   everywhere) | 'fp32' (SIMT parity mode, logits within 1e-4)
      tokenizer    any ``str -> LongTensor[1,77]`` callable.  Default: with REAL weights (``pretrained_path`` / ``state_dict``) the
                   reference's own ``open_clip.get_tokenizer(model_name)`` (clip_wrapper.py:27) -- and a RuntimeError if open_clip is
                   not importable, because token ids from any other vocabulary make the class prompts meaningless; with random-init
                   (synthetic / benchmark) weights the dependency-free SyntheticTokenizer.  Pass ``tokenizer="synthetic"`` to force it.
    """

    def __init__(self, model_name="ViT-B-32", pretrained_path=None, device="cuda", *, state_dict=None, seed=0,
                 attribution="literal", dtype="mixed", tokenizer=None):
        super().__init__()
        if attribution not in ("literal", "intended"):
            raise ValueError(f"Unknown attribution mode: {attribution}")
        self.device = device
        self.attribution = attribution
        self.dtype = dtype
        cfg = get_model_config(model_name)
        self.model = EngineCLIP(cfg).to(device)                      # clip_wrapper.py:13,16
        # open_clip's inference transform (Resize-BICUBIC / CenterCrop / ToTensor / Normalize) on the device (clip_wrapper.py:13)
        self.preprocess = GpuPreprocess(cfg.image_size, device=device) if torch.device(device).type == "cuda" else None
        if pretrained_path is not None and state_dict is None:       # clip_wrapper.py:14
            state_dict = torch.load(pretrained_path, map_location=device)
        self.engine = Engine(cfg, dtype=dtype, device=device)
        self.model._engine = self.engine
        if state_dict is not None:
            self.model.load_state_dict(state_dict, strict=True)      # clip_wrapper.py:15
        else:
            random_init_(self.model, seed)
        self.model.eval()
        for p in self.model.parameters():                            # clip_wrapper.py:19-20
            p.requires_grad = False
        self.sync_engine_weights()
        self.model.register_load_state_dict_post_hook(lambda module, incompatible: self.sync_engine_weights())
        self.attention_maps = []                                     # clip_wrapper.py:23
        self.tokenizer = self._pick_tokenizer(tokenizer, model_name, cfg, real_weights=state_dict is not None)   # clip_wrapper.py:27

    @staticmethod
    def _pick_tokenizer(tokenizer, model_name, cfg, real_weights):
        if tokenizer == "synthetic":
            return SyntheticTokenizer(cfg.context_length)
        if tokenizer is not None:
            return tokenizer
        if not real_weights:
            return SyntheticTokenizer(cfg.context_length)            # random-init weights: any consistent id assignment will do
        try:
            import open_clip                                         # the reference's tokenizer source (clip_wrapper.py:5,27)
        except ImportError as e:
            raise RuntimeError(
                "CLIPWrapper was given real weights but no tokenizer, and open_clip (whose BPE vocabulary those weights were trained "
                "with) is not importable. Pass tokenizer=<str -> LongTensor[1,77] callable> (e.g. open_clip.get_tokenizer(model_name)), "
                "or tokenizer='synthetic' if hashed token ids are really what you want.") from e
        return open_clip.get_tokenizer(model_name)

    def sync_engine_weights(self):
        self.engine.load_state_dict(self.model.state_dict())

    def reset(self):                                                 # clip_wrapper.py:42-44
        self.attention_maps.clear()

    def encode_image(self, image_tensor):                            # clip_wrapper.py:46-47
        return self.model.encode_image(image_tensor)

    def encode_text(self, token_tensor):                             # clip_wrapper.py:49-51
        self.reset()
        return self.model.encode_text(token_tensor)

    def get_attention_map(self):                                     # clip_wrapper.py:53-59
        if len(self.attention_maps) == 0:
            return None
        return self.attention_maps[-1]

    def get_tokenizer(self):                                         # clip_wrapper.py:61-62
        return self.tokenizer

    def get_preprocess(self):                                        # clip_wrapper.py:64-65
        return self.preprocess
