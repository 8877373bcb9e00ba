"""CPU restatement of the reference's own hot-path modules (TEST INFRASTRUCTURE).

Each class/function cites the reference lines it follows.  ``forward_as_written``
keeps the reference's B x n_cls double loop (2*B*n_cls text-transformer calls);
``forward_dedup`` is the same arithmetic with the loops hoisted (the text side
does not depend on the sample index b — SURVEY fact 8) and is what the CUDA
path is compared against at full size.  ``oracle/make_goldens.py`` checks both
against the unmodified reference modules imported from /root/reference.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class OraclePromptLearner(nn.Module):
    """models/prompt_learner.py:5-70 (device default made a parameter; nothing else changes)."""

    def __init__(self, class_names, clip_model, prompt_len=5, class_specific=True,
                 use_init_prompt=True, device="cpu", ctx_seed=None):
        super().__init__()
        self.prompt_len = prompt_len
        self.class_specific = class_specific
        self.ctx_dim = clip_model.model.token_embedding.embedding_dim      # :11
        self.tokenizer = clip_model.get_tokenizer()                        # :12
        self.token_embedding = clip_model.model.token_embedding            # :13
        self.device = device
        self.use_init_prompt = use_init_prompt
        self.context_bank = nn.ParameterDict()                             # :18
        self.token_bank = {}                                               # :19
        self._gen = None if ctx_seed is None else torch.Generator().manual_seed(ctx_seed)
        for name in class_names:                                           # :23-24
            self.add_class_prompt(name)

    def add_class_prompt(self, class_name):                                # :26-43
        if class_name in self.context_bank:
            return
        with torch.no_grad():
            text = f"a photo of a {class_name}"
            tokenized = self.tokenizer(text).to(self.device)               # [1, 77]
            token_emb = self.token_embedding(tokenized.unsqueeze(0)).squeeze(0)   # [1, 77, D]
            self.token_bank[class_name] = token_emb
            if self.use_init_prompt and token_emb.shape[0] >= 5 + self.prompt_len:   # never true: shape[0] == 1
                ctx_init = token_emb[5:5 + self.prompt_len].clone()
            elif self._gen is None:
                ctx_init = torch.randn(self.prompt_len, self.ctx_dim).to(self.device)   # :41
            else:
                ctx_init = torch.randn(self.prompt_len, self.ctx_dim, generator=self._gen).to(self.device)
        self.context_bank[class_name] = nn.Parameter(ctx_init)

    def forward(self):                                                     # :45-66
        prompts = []
        for cls in self.context_bank:
            ctx = self.context_bank[cls].unsqueeze(0)
            token = self.token_bank[cls]
            if token.dim() == 2:
                token = token.unsqueeze(0)
            elif token.dim() == 4:
                token = token.squeeze(0)
            elif token.dim() != 3:
                raise ValueError(f"Unexpected token shape: {token.shape}")
            prompts.append(torch.cat([ctx, token], dim=1))
        return torch.cat(prompts, dim=0)

    @property
    def n_cls(self):
        return len(self.context_bank)


def attribution_monitor(attn_map, prompt_len, normalize=True):
    """models/attribution_monitor.py:17-36."""
    B, T, _ = attn_map.shape
    raw_score = attn_map[:, :prompt_len, T - 1]
    return F.softmax(raw_score, dim=-1) if normalize else raw_score


def prompt_adjustor_scale(prompt_embed, attribution_score):
    """models/prompt_adjustor.py:27-36, method='scale'."""
    return prompt_embed * attribution_score.unsqueeze(-1)


class OraclePromptAdjustor(nn.Module):
    """models/prompt_adjustor.py:6-47, all three methods (the reference raises ValueError for an unknown method at the first
    forward, :47; here at construction)."""

    def __init__(self, method="scale"):
        super().__init__()
        self.method = method
        if method == "gate":                                                   # :13-19
            self.gate_net = nn.Sequential(nn.Linear(1, 64), nn.ReLU(), nn.Linear(64, 1), nn.Sigmoid())
        elif method == "residual":                                             # :20-25
            self.residual_net = nn.Sequential(nn.Linear(1, 64), nn.ReLU(), nn.Linear(64, 512))
        elif method != "scale":
            raise ValueError(f"Unknown method: {method}")

    def forward(self, prompt_embed, attribution_score):                        # :27-44
        a = attribution_score.unsqueeze(-1)
        if self.method == "scale":
            return prompt_embed * a
        if self.method == "gate":
            return prompt_embed * self.gate_net(a)
        return prompt_embed + self.residual_net(a)


class OracleFullModel(nn.Module):
    """models/model_wrapper.py:12-100."""

    def __init__(self, class_names, clip_wrapper, prompt_len=5, attr_lambda=1.0, stab_lambda=0.1,
                 adjustor_method="scale", class_specific=False, ctx_seed=None):
        super().__init__()
        self.clip = clip_wrapper
        self.class_names = class_names
        self.prompt_learner = OraclePromptLearner(class_names, clip_wrapper, prompt_len, class_specific,
                                                  device=clip_wrapper.device, ctx_seed=ctx_seed)
        self.n_cls = len(class_names)
        self.prompt_len = prompt_len
        self.prompt_adjustor = OraclePromptAdjustor(adjustor_method)         # :22 (created after the prompt learner: RNG order)
        self.attr_lambda, self.stab_lambda = attr_lambda, stab_lambda
        self.logit_scale = nn.Parameter(torch.ones([]) * torch.log(torch.tensor(1 / 0.07)))   # :26

    # -- the reference schedule, loop for loop ---------------------------------------------
    def forward_as_written(self, images, labels=None):                     # :28-100
        B = images.size(0)
        raw_prompt = self.prompt_learner()                                 # :32
        P = self.prompt_learner.prompt_len
        context_prompt, class_tokens = raw_prompt[:, :P, :], raw_prompt[:, P:, :]   # :34-35
        image_feat = self.clip.encode_image(images)                        # :40
        image_feat = image_feat / image_feat.norm(dim=-1, keepdim=True)    # :41
        logits = []
        for i, _ in enumerate(list(self.prompt_learner.context_bank.keys())):   # :47-48
            ctx = context_prompt[i].unsqueeze(0).expand(B, -1, -1)
            cls = class_tokens[i].unsqueeze(0).expand(B, -1, -1)
            full_prompt = torch.cat([ctx, cls], dim=1)                     # :51
            attributions = []
            for b in range(B):                                             # :55-63
                self.clip.reset()
                _ = self.clip.model.transformer(full_prompt[b].unsqueeze(0))
                attn_map = self.clip.get_attention_map()
                if attn_map.dim() == 2:
                    attn_map = attn_map.unsqueeze(0)
                attributions.append(attribution_monitor(attn_map, P))
            attribution = torch.cat(attributions, dim=0)                   # :65
            adjusted_ctx = self.prompt_adjustor(ctx, attribution)          # :68
            adjusted_prompt = torch.cat([adjusted_ctx, cls], dim=1)        # :69
            text_feat = self.clip.model.transformer(adjusted_prompt)       # :72
            text_feat = text_feat[torch.arange(B), -1, :]                  # :73
            text_feat = text_feat @ self.clip.model.text_projection        # :74
            text_feat = text_feat / text_feat.norm(dim=-1, keepdim=True)   # :75
            logits.append(self.logit_scale.exp() * (image_feat * text_feat).sum(dim=-1, keepdim=True))   # :79
        logits = torch.cat(logits, dim=1)                                  # :83
        outputs = {"logits": logits}
        if labels is not None:                                             # :90-93
            loss_cls = F.cross_entropy(logits, labels)
            outputs.update({"loss": loss_cls, "loss_cls": loss_cls})
        return outputs

    forward = forward_as_written

    # -- same arithmetic, loops hoisted -----------------------------------------------------
    def text_attribution(self, raw_prompt):
        """Rows A7/A8: one un-adjusted pass over [C,T,D]; returns (raw [C,P'], attribution [C,P'])."""
        P = self.prompt_len
        self.clip.reset()
        with torch.no_grad():
            _ = self.clip.model.transformer(raw_prompt.detach())
        attn_map = self.clip.get_attention_map()          # literal: [C, D]; intended: [C, T, T]
        if attn_map.dim() == 2:                           # per-sample [D] -> unsqueeze(0) -> [1,1,D] in the loop form
            attn_map = attn_map.unsqueeze(1)              # [C, 1, D]: T'=1 per class
        T = attn_map.shape[1]
        raw = attn_map[:, :P, T - 1]
        return raw, F.softmax(raw, dim=-1)

    def text_features(self, raw_prompt, attribution):
        """Rows A9/A10: adjust ctx, feature pass, last-position pool, projection, L2-norm -> [C,E]."""
        P = self.prompt_len
        adjusted = torch.cat([self.prompt_adjustor(raw_prompt[:, :P, :], attribution), raw_prompt[:, P:, :]], dim=1)
        x = self.clip.model.transformer(adjusted)
        feat = x[:, -1, :] @ self.clip.model.text_projection
        return feat / feat.norm(dim=-1, keepdim=True)

    def forward_dedup(self, images, labels=None, return_aux=False):
        raw_prompt = self.prompt_learner()
        with torch.no_grad():
            image_feat = self.clip.encode_image(images)
            image_feat = image_feat / image_feat.norm(dim=-1, keepdim=True)
        raw, attribution = self.text_attribution(raw_prompt)
        text_feat = self.text_features(raw_prompt, attribution.detach())
        logits = self.logit_scale.exp() * image_feat @ text_feat.t()
        outputs = {"logits": logits}
        if labels is not None:
            loss_cls = F.cross_entropy(logits, labels)
            outputs.update({"loss": loss_cls, "loss_cls": loss_cls})
        if return_aux:
            outputs.update({"image_feat": image_feat, "text_feat": text_feat,
                            "attr_raw": raw, "attribution": attribution})
        return outputs


# ----------------------------------------------------------------------------------------
# seeded synthetic inputs shared by the oracle, the tests and bench.py
# ----------------------------------------------------------------------------------------
def class_names(n_cls: int):
    return [f"class_{i:03d}" for i in range(n_cls)]


def synthetic_images(batch: int, image_size: int, seed: int = 1):
    return torch.randn(batch, 3, image_size, image_size, generator=torch.Generator().manual_seed(seed))


def synthetic_labels(batch: int, n_cls: int, seed: int = 2):
    return torch.randint(0, n_cls, (batch,), generator=torch.Generator().manual_seed(seed))
