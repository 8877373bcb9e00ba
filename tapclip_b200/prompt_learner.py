"""Drop-in for ``models/prompt_learner.py`` (PromptLearner, lines 5-70) with an engine-friendly layout.

Same surface: ``context_bank`` (nn.ParameterDict, insertion order = class index), ``token_bank`` (plain dict of
frozen ``[1,77,D]`` token embeddings), ``add_class_prompt(name)`` usable after construction, ``forward()`` ->
``[n_cls, P+77, D]``, ``n_cls``.  State-dict keys stay ``context_bank.<class>``.

Layout: every ``context_bank[cls]`` Parameter is a view into ONE contiguous fp32 bank ``[capacity, P, D]`` and every
``token_bank[cls]`` a view into ``[capacity, 77, D]``, so the engine reads all class prompts through two pointers
(no per-class ``cat``: prompt_learner.py:62,65 launches n_cls+1 kernels per forward) and optimizers update the bank
in place.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class PromptLearner(nn.Module):
    def __init__(self, class_names, clip_model, prompt_len=5, class_specific=True,
                 use_init_prompt=True, device="cuda"):
        super().__init__()
        self.prompt_len = prompt_len
        self.class_specific = class_specific              # stored, unused — as in the reference
        self.ctx_dim = clip_model.model.token_embedding.embedding_dim
        self.tokenizer = clip_model.get_tokenizer()
        self.token_embedding = clip_model.model.token_embedding     # frozen
        self.device = device
        self.use_init_prompt = use_init_prompt
        self.context_bank = nn.ParameterDict()
        self.token_bank = {}
        self._ctx_flat = None          # [capacity, P, D]
        self._tok_flat = None          # [capacity, L, D]
        for name in class_names:
            self.add_class_prompt(name)

    # ---- storage ----------------------------------------------------------------------------------
    def _reserve(self, n, tok_len):
        cap = 0 if self._ctx_flat is None else self._ctx_flat.shape[0]
        if n <= cap:
            return
        new_cap = max(n, 2 * cap, 8)
        dev = self.token_embedding.weight.device
        ctx = torch.zeros(new_cap, self.prompt_len, self.ctx_dim, device=dev)
        tok = torch.zeros(new_cap, tok_len, self.ctx_dim, device=dev)
        names = list(self.context_bank.keys())
        if names:
            ctx[: len(names)] = torch.stack([self.context_bank[k].detach() for k in names])
            tok[: len(names)] = torch.cat([self.token_bank[k] for k in names], dim=0)
        self._ctx_flat, self._tok_flat = ctx, tok
        self._rebind()

    def _rebind(self):
        """Point every Parameter / token entry at its row of the flat banks (Parameter identity is kept)."""
        for i, k in enumerate(self.context_bank.keys()):
            self.context_bank[k].data = self._ctx_flat[i]
            self.token_bank[k] = self._tok_flat[i: i + 1]

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        names = list(self.context_bank.keys())
        if names and self._ctx_flat is not None:
            first = self.context_bank[names[0]]
            if first.data_ptr() != self._ctx_flat.data_ptr() or first.device != self._ctx_flat.device:
                ctx = torch.stack([self.context_bank[k].detach() for k in names])
                tok = torch.cat([self.token_bank[k].to(ctx.device) for k in names], dim=0)
                self._ctx_flat = torch.zeros(self._ctx_flat.shape, device=ctx.device)
                self._tok_flat = torch.zeros(self._tok_flat.shape, device=ctx.device)
                self._ctx_flat[: len(names)] = ctx
                self._tok_flat[: len(names)] = tok
                self._rebind()
        return out

    # ---- reference API ---------------------------------------------------------------------------------
    def add_class_prompt(self, class_name):                        # prompt_learner.py:26-43
        if class_name in self.context_bank:
            return
        with torch.no_grad():
            text = f"a photo of a {class_name}"
            tokenized = self.tokenizer(text).to(self.token_embedding.weight.device)        # [1, 77]
            token_emb = self.token_embedding(tokenized.unsqueeze(0)).squeeze(0)            # [1, 77, D]
            if token_emb.dim() != 3 or token_emb.shape[0] != 1:
                raise ValueError(f"Unexpected token shape: {token_emb.shape}")
            if self.use_init_prompt and token_emb.shape[0] >= 5 + self.prompt_len:         # never true (shape[0]==1), kept
                ctx_init = token_emb[5:5 + self.prompt_len].clone()
            else:
                ctx_init = torch.randn(self.prompt_len, self.ctx_dim).to(token_emb.device)  # CPU RNG, as the reference
        i = len(self.context_bank)
        self._reserve(i + 1, token_emb.shape[1])
        self._ctx_flat[i].copy_(ctx_init)
        self._tok_flat[i].copy_(token_emb[0])
        self.token_bank[class_name] = self._tok_flat[i: i + 1]
        self.context_bank[class_name] = nn.Parameter(self._ctx_flat[i])

    def forward(self):                                             # prompt_learner.py:45-66
        n = self.n_cls
        return torch.cat([self.flat_ctx(), self._tok_flat[:n]], dim=1)

    @property
    def n_cls(self):
        return len(self.context_bank)

    # ---- engine-facing views -------------------------------------------------------------------------------
    def flat_ctx(self):
        """[n_cls, P, D] view of the bank; re-packs if a Parameter was re-pointed behind our back."""
        n = self.n_cls
        for i, k in enumerate(self.context_bank.keys()):
            p = self.context_bank[k]
            if p.data_ptr() != self._ctx_flat[i].data_ptr():
                with torch.no_grad():
                    self._ctx_flat[i].copy_(p.detach())
                p.data = self._ctx_flat[i]
        return self._ctx_flat[:n]

    def flat_tok(self):
        return self._tok_flat[: self.n_cls]
