"""Fused AdamW over the flat ctx bank (replaces torch.optim.AdamW of train.py:65-67,105
for the prompt parameters): one kernel over ``[C,P,D]`` instead of a foreach over n_cls tensors.

Semantics are torch.optim.AdamW's (decoupled weight decay, bias-corrected moments, eps outside the sqrt).
"""
from __future__ import annotations

import torch


class FusedAdamW:
    def __init__(self, model, lr=2e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.model = model                      # a tapclip_b200.FullModel
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        self.exp_avg = self.exp_avg_sq = None

    def _grad_flat(self):
        pl = self.model.prompt_learner
        params = list(pl.context_bank.values())
        g0 = params[0].grad
        if g0 is None:
            return None
        n = len(params)
        stride = g0.numel()
        # gradients produced by FullModel.backward are rows of one [C,P,D] buffer: use it without a copy
        base = g0.data_ptr()
        if all(p.grad is not None and p.grad.is_contiguous() and p.grad.data_ptr() == base + i * stride * 4 for i, p in enumerate(params)):
            return torch.as_strided(g0, (n * stride,), (1,)) if g0.untyped_storage().nbytes() - (g0.storage_offset() * 4) >= n * stride * 4 else None
        return None

    def zero_grad(self, set_to_none=True):
        for p in self.model.prompt_learner.context_bank.values():
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self):
        pl = self.model.prompt_learner
        bank = pl.flat_ctx()
        flat = bank.view(-1)
        g = self._grad_flat()
        if g is None:
            grads = [p.grad for p in pl.context_bank.values()]
            if any(x is None for x in grads):
                return
            g = torch.stack(grads).reshape(-1)
        if self.exp_avg is None or self.exp_avg.numel() != flat.numel():
            old = self.exp_avg
            self.exp_avg = torch.zeros_like(flat)
            self.exp_avg_sq_new = torch.zeros_like(flat)
            if old is not None:                 # classes were added: keep the moments of the existing ones
                k = min(old.numel(), flat.numel())
                self.exp_avg[:k] = old[:k]
                self.exp_avg_sq_new[:k] = self.exp_avg_sq[:k]
            self.exp_avg_sq = self.exp_avg_sq_new
            del self.exp_avg_sq_new
        self.step_count += 1
        self.model.clip.engine.adamw_step(flat, g.contiguous(), self.exp_avg, self.exp_avg_sq, self.lr, self.betas, self.eps,
                                          self.weight_decay, self.step_count)
        for p in pl.context_bank.values():      # raw-pointer update: make autograd / the text-feature cache see it
            torch.autograd.graph.increment_version(p)
