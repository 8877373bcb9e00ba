"""Drop-in for ``models/prompt_adjustor.py`` (PromptAdjustor, lines 6-47).

Only ``method='scale'`` (lines 35-36) is on the hot path — it is the only method any reference script
selects, and inside ``FullModel`` it is fused into the ctx-splice kernel (K4) so the scaled context is
never materialised per sample.  'gate' / 'residual' (lines 13-25, 38-44) are never selected, never
given to the optimizer (train.py:65-67) and are out of scope.
"""
from __future__ import annotations

import torch.nn as nn


class PromptAdjustor(nn.Module):
    def __init__(self, method="scale"):
        super().__init__()
        if method in ("gate", "residual"):
            raise NotImplementedError(f"PromptAdjustor method {method!r} is outside the B200 hot path (only 'scale' is used by the reference scripts)")
        if method != "scale":
            raise ValueError(f"Unknown method: {method}")      # prompt_adjustor.py:47
        self.method = method

    def forward(self, prompt_embed, attribution_score):
        """prompt_embed [B,P,D] * attribution_score [B,P or 1] (API shim; the engine fuses this into the splice)."""
        return prompt_embed * attribution_score.unsqueeze(-1)
