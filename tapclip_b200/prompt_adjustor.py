"""Drop-in for ``models/prompt_adjustor.py`` (PromptAdjustor, lines 6-47).

``method='scale'`` (lines 35-36) is the only method any reference script selects; inside ``FullModel`` it is fused into
the ctx-splice kernel (K4), so the scaled context is never materialised per sample.

``'gate'`` / ``'residual'`` (lines 13-25, 38-44) keep the reference's module structure (same
sub-module and state-dict names: ``gate_net.{0,2}.*`` / ``residual_net.{0,2}.*``).  Their networks act on the
``[C, P]`` attribution scores only (a 1 -> 64 -> 1 or 1 -> 64 -> D perceptron per ctx token, ~1e5 flop), between the
engine's attribution pass and its feature pass; ``FullModel`` runs them as ordinary autograd ops on the de-duplicated
``[C, P, D]`` context bank, so their parameters receive gradients exactly as in the reference.
"""
from __future__ import annotations

import torch.nn as nn


class PromptAdjustor(nn.Module):
    def __init__(self, method="scale", dim=512):
        super().__init__()
        self.method = method
        if method == "gate":                                   # prompt_adjustor.py:13-19
            self.gate_net = nn.Sequential(nn.Linear(1, 64), nn.ReLU(), nn.Linear(64, 1), nn.Sigmoid())
        elif method == "residual":                             # prompt_adjustor.py:20-25 (hard-codes 512 = ViT-B text width)
            self.residual_net = nn.Sequential(nn.Linear(1, 64), nn.ReLU(), nn.Linear(64, dim))
        elif method != "scale":
            raise ValueError(f"Unknown method: {method}")      # prompt_adjustor.py:47 (raised at first use there)

    def forward(self, prompt_embed, attribution_score):
        """prompt_embed [B,P,D], attribution_score [B,P] (or [B,1]: the literal-mode attribution, broadcast over P)."""
        a = attribution_score.unsqueeze(-1)                    # [B, P|1, 1]
        if self.method == "scale":
            return prompt_embed * a
        if self.method == "gate":
            return prompt_embed * self.gate_net(a)
        return prompt_embed + self.residual_net(a)
