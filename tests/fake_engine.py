"""TEST DOUBLE for tapclip_b200.engine.Engine: the same method surface, backed by the CPU oracle's torch modules.

Lets the host-side logic of tapclip_b200 (FullModel's autograd Function, flat ctx bank, class sharding and the
collectives around it) run on CPU / gloo.  It lives in tests/ and is never importable from the product package.
"""
import torch
import torch.nn.functional as F


class FakeEngine:
    def __init__(self, oracle_wrapper):
        self.ow = oracle_wrapper
        self.launch_count = 0
        self._saved = None
        self.last_forward_token = 0          # same contract as the real engine: one saved forward, identified by a token
        self.weight_generation = 0

    def encode_image(self, images, want_cls_rows=False):
        with torch.no_grad():
            return self.ow.model.encode_image(images)

    def _attribution(self, raw_prompt, P, mode):
        if mode == "literal":
            n = raw_prompt.shape[0]
            return torch.ones(n, 1), None
        blk = self.ow.model.transformer.resblocks[-1].attn
        saved, grabbed = blk.emit, []
        blk.emit = "weights"
        h = blk.register_forward_hook(lambda m, i, o: grabbed.append(o[0].detach().mean(dim=1)))
        try:
            with torch.no_grad():
                self.ow.model.transformer(raw_prompt)
        finally:
            h.remove()
            blk.emit = saved
        amap = grabbed[-1]
        raw = amap[:, :P, amap.shape[1] - 1].clone()
        return F.softmax(raw, dim=-1), raw

    def text_attribution(self, ctx, tok):
        return self._attribution(torch.cat([ctx, tok], dim=1), ctx.shape[1], "intended")

    def text_forward(self, ctx, tok, mode, save_for_backward):
        P = ctx.shape[1]
        attr, raw = self._attribution(torch.cat([ctx, tok], dim=1), P, mode)
        leaf = ctx.detach().clone().requires_grad_(save_for_backward)
        with torch.set_grad_enabled(save_for_backward):
            x = self.ow.model.transformer(torch.cat([leaf * attr.unsqueeze(-1), tok], dim=1))
            feat = x[:, -1, :] @ self.ow.model.text_projection
            feat = feat / feat.norm(dim=-1, keepdim=True)
        self._saved = (leaf, feat) if save_for_backward else None
        self.last_forward_token = self.last_forward_token + 1 if save_for_backward else 0
        return feat.detach(), attr, raw

    def logits(self, img_feat, text_feat, logit_scale, labels=None, inv_batch_total=None):
        img_norm = img_feat / img_feat.norm(dim=-1, keepdim=True)
        logits = logit_scale.detach().exp() * img_norm @ text_feat.t()
        if labels is None:
            return logits, None, None, img_norm
        inv = inv_batch_total if inv_batch_total is not None else 1.0 / logits.shape[0]
        lse = torch.logsumexp(logits, dim=1)
        loss = ((lse - logits.gather(1, labels[:, None]).squeeze(1)).sum() * inv)
        dl = (torch.softmax(logits, 1) - F.one_hot(labels, logits.shape[1]).float()) * inv
        return logits, loss, dl, img_norm

    def logits_backward(self, dlogits, logits, img_norm, logit_scale):
        d_text = logit_scale.detach().exp() * dlogits.t() @ img_norm
        return d_text, (dlogits * logits).sum()

    def text_backward(self, d_text_feat, n_cls, prompt_len, token=0):
        if self._saved is None or (token and token != self.last_forward_token):
            raise RuntimeError("stale backward: the saved activations were replaced by a later forward")
        leaf, feat = self._saved
        (g,) = torch.autograd.grad(feat, leaf, d_text_feat)
        return g


class FakeWrapper(torch.nn.Module):
    """Duck-typed stand-in for tapclip_b200.CLIPWrapper on CPU (oracle CLIP + FakeEngine)."""

    def __init__(self, oracle_wrapper, attribution):
        super().__init__()
        self.model = oracle_wrapper.model
        self.device = "cpu"
        self.attribution = attribution
        self.engine = FakeEngine(oracle_wrapper)
        self.attention_maps = []
        self.tokenizer = oracle_wrapper.get_tokenizer()

    def get_tokenizer(self):
        return self.tokenizer

    def get_attention_map(self):
        return self.attention_maps[-1] if self.attention_maps else None
