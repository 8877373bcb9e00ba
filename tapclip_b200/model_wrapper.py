"""Drop-in for ``models/model_wrapper.py`` (FullModel, lines 12-100) on B200.

Same constructor, same ``forward(images, labels=None) -> {"logits", "loss", "loss_cls"}``, same parameter /
state-dict names; ``loss.backward()`` fills ``.grad`` of every ``prompt_learner.context_bank[cls]`` and of
``logit_scale``.  The schedule underneath is the de-duplicated one:

    reference (model_wrapper.py:47-81)             here
    ---------------------------------------------  ---------------------------------------------------
    B*n_cls batch-1 text passes for attribution    ONE [C,T,D] attribution pass (probe epilogue, K2+K3)
    n_cls batch-B text passes for features         ONE [C,T,D] feature pass on [ctx*a | tok]   (K4, K1, K2)
    n_cls x (mul, sum, exp, cat)                   one cosine-logit kernel + fused cross-entropy

Multi-GPU (one process per GPU, ``torch.distributed`` initialised by the caller): images are data-parallel,
class prompts are sharded across ranks, text features are all-gathered, the text-feature gradient is
all-reduced and the ctx gradients are all-gathered so that every replica holds the full, identical gradient
(SURVEY.md 8e).  Pass ``distributed=True`` (default: auto when a process group exists).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .attribution_monitor import AttributionMonitor
from .parallel import ClassSharding, TextGather, all_gather_rows, all_reduce_sum_
from .prompt_adjustor import PromptAdjustor
from .prompt_learner import PromptLearner


class _TapClipFunction(torch.autograd.Function):
    """images, labels, ctx bank, token bank -> logits (+ loss); backward -> ctx / logit_scale gradients."""

    @staticmethod
    def forward(ctx, model, images, labels, need_grad, logit_scale, *ctx_params):
        # ctx_params: the n_cls context Parameters ('scale': the engine reads them through the flat bank), or ONE adjusted
        # [C,P,D] tensor produced by the 'gate' / 'residual' adjustor (then the engine runs its literal feature pass on it)
        eng = model.clip.engine
        pl = model.prompt_learner
        mode = model.clip.attribution
        shard = model._sharding()
        n_cls, P = pl.n_cls, pl.prompt_len
        lo, hi = shard.bounds(n_cls)
        adjusted = ctx_params[0].detach().contiguous() if model._adjusted_path() else None
        ctx.adjusted = adjusted is not None

        # The two towers are independent until the logit contraction: the (small, launch-bound) text passes run on a
        # side stream and fill the image tower's kernel tails; joined before the logits.
        # multi-GPU: the text-feature all-gather is either fused into the head kernel (peer stores into every rank's symmetric buffer,
        # parallel.TextGather; used whenever the features are recomputed per step) or one NCCL all-gather (cached eval features, gloo)
        tg = model._text_gather(shard, n_cls, images.device, need_grad)
        epoch = tg.next_epoch() if tg is not None else 0
        gather = (lo, epoch) if tg is not None else None

        def text_side():
            local, attr_l, raw_l = model._text_features(lo, hi, need_grad, adjusted, gather)   # rows A6-A10, this rank's classes
            feat = tg.slot(epoch) if tg is not None else all_gather_rows(local, shard, n_cls)   # [C, E]
            return local, attr_l, raw_l, feat

        side = model._side_stream(images.device)
        if side is not None:
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                text_local, attr_local, raw_local, text_feat = text_side()        # the collective hides behind the image tower too
            img_feat = model._encode_image(images)                                # row A4 (+ A-ext probes)
            main.wait_stream(side)
            for t in (text_local, attr_local, raw_local, text_feat):
                if t is not None:
                    t.record_stream(main)
        else:
            img_feat = model._encode_image(images)
            text_local, attr_local, raw_local, text_feat = text_side()
        ctx.reduced = None
        ge = {"gather_epoch": epoch} if epoch else {}
        if labels is not None:
            b_total = shard.global_batch(images.shape[0])
            logits, loss, dlogits_ce, img_norm = eng.logits(img_feat, text_feat, logit_scale, labels, 1.0 / b_total, **ge)
            if shard.world > 1 and (need_grad or logit_scale.requires_grad):
                # ONE collective for the whole logit head: the CE gradient is already known here (the fused logits kernel emits
                # dloss/dlogits), so d T^ [C,E], d logit_scale and the loss are reduced together; backward scales the cached
                # result by the incoming d loss (they are linear in it) instead of issuing two more all-reduces
                d_text, d_scale = eng.logits_backward(dlogits_ce, logits, img_norm, logit_scale)
                packed = torch.cat([d_text.reshape(-1), d_scale.reshape(1), loss.reshape(1)])
                all_reduce_sum_(packed, shard)
                n = d_text.numel()
                ctx.reduced = (packed[:n].view_as(d_text), packed[n])
                loss = packed[n + 1].clone()
            elif shard.world > 1:
                loss = all_reduce_sum_(loss.clone(), shard)
        else:
            logits, loss, dlogits_ce, img_norm = eng.logits(img_feat, text_feat, logit_scale, **ge)
        if adjusted is None:
            model.clip.attention_maps[:] = [raw_local if raw_local is not None else attr_local]   # compact probe, see clip_wrapper.py
            model.last_attribution = attr_local
        ctx.model, ctx.shard, ctx.dims = model, shard, (n_cls, P, lo, hi)
        ctx.need_grad = need_grad
        # the engine keeps ONE set of saved text activations: remember which forward wrote them, so a backward that comes after
        # another forward on the same CLIPWrapper (two FullModels sharing it, two forwards before one backward) fails loudly
        ctx.text_token = eng.last_forward_token if need_grad else 0
        ctx.save_for_backward(logits, dlogits_ce if dlogits_ce is not None else logits.new_empty(0), img_norm, logit_scale)
        ctx.has_ce = dlogits_ce is not None
        ctx.set_materialize_grads(False)
        if loss is None:
            loss = logits.new_zeros(())
        return logits, loss

    @staticmethod
    def backward(ctx, g_logits, g_loss):
        model, shard = ctx.model, ctx.shard
        eng = model.clip.engine
        n_cls, P, lo, hi = ctx.dims
        logits, dlogits_ce, img_norm, logit_scale = ctx.saved_tensors
        n_in = 5 + (1 if ctx.adjusted else n_cls)
        if ctx.reduced is not None and g_logits is None and g_loss is not None:
            # the usual `loss.backward()`: the reduced gradients of the head were computed (and all-reduced) in forward
            d_text, d_scale = ctx.reduced[0] * g_loss, ctx.reduced[1] * g_loss
        else:
            dl = None
            if ctx.has_ce and g_loss is not None:
                dl = dlogits_ce * g_loss
            if g_logits is not None:
                dl = g_logits.contiguous() if dl is None else dl + g_logits
            if dl is None:
                return (None,) * n_in
            d_text, d_scale = eng.logits_backward(dl.contiguous(), logits, img_norm, logit_scale)
            if shard.world > 1:                                                   # one collective for both
                packed = torch.cat([d_text.reshape(-1), d_scale.reshape(1)])
                all_reduce_sum_(packed, shard)
                d_text, d_scale = packed[:-1].view_as(d_text), packed[-1]
        grads = [None] * n_cls
        if ctx.need_grad:
            d_local = d_text[lo:hi].contiguous()
            dctx_local = eng.text_backward(d_local, hi - lo, P, token=ctx.text_token)   # row A13
            dctx = all_gather_rows(dctx_local.view(hi - lo, -1), shard, n_cls).view(n_cls, P, -1)
            grads = [dctx] if ctx.adjusted else list(dctx.unbind(0))
        elif ctx.adjusted:
            grads = [None]
        return (None, None, None, None, d_scale.reshape(logit_scale.shape) if logit_scale.requires_grad else None, *grads)


class FullModel(nn.Module):
    """models/model_wrapper.py:12-100."""

    def __init__(self, class_names, clip_wrapper, prompt_len=5, attr_lambda=1.0, stab_lambda=0.1,
                 adjustor_method="scale", class_specific=False, *, distributed=None, cache_text_features=True,
                 overlap_towers=True, image_attribution=None):
        super().__init__()
        self.clip = clip_wrapper
        self.class_names = class_names
        self.prompt_learner = PromptLearner(class_names, clip_wrapper, prompt_len, class_specific,
                                            device=clip_wrapper.device)
        self.n_cls = len(class_names)
        self.attribution_monitor = AttributionMonitor(prompt_len)
        self.prompt_adjustor = PromptAdjustor(method=adjustor_method, dim=clip_wrapper.model.token_embedding.embedding_dim)
        self.prompt_adjustor.to(clip_wrapper.model.text_projection.device)
        self.attr_lambda = attr_lambda          # stored, never read — as in the reference (model_wrapper.py:24-25)
        self.stab_lambda = stab_lambda
        dev = clip_wrapper.model.text_projection.device
        self.logit_scale = nn.Parameter((torch.ones([]) * torch.log(torch.tensor(1 / 0.07))).to(dev))
        self.distributed = distributed
        self.cache_text_features = cache_text_features
        self._text_cache = None
        self.last_attribution = None
        self.overlap_towers = overlap_towers
        self._side = None
        self.fused_gather = True          # multi-GPU: text-feature all-gather fused into the head kernel when available (TextGather)
        self._tg = None
        # north-star extension (SURVEY 8a row A-ext, not in the reference): None | 'cls' (per-layer, per-head CLS-row
        # attention probabilities [B,L,H,N]) | 'rollout' (additionally the attention-rollout map [B,N-1]); emitted by the
        # same image pass that produces the features, returned in the forward() dict
        if image_attribution not in (None, "cls", "rollout"):
            raise ValueError(f"Unknown image_attribution: {image_attribution}")
        self.image_attribution = image_attribution
        self.last_image_attribution = None

    @property
    def prompt_len(self):
        return self.prompt_learner.prompt_len

    def _side_stream(self, device):
        """Side stream for the text tower (None = run both towers on the caller's stream)."""
        if not self.overlap_towers or device.type != "cuda":
            return None
        if self._side is None:
            # TAPCLIP_TEXT_PRIO=1: the text tower's (small, dependent) kernels get the SMs first; the image tower's GEMMs draw their
            # tiles dynamically (gemm_tc.cu) and take whatever is free
            prio = -1 if os.environ.get("TAPCLIP_TEXT_PRIO", "0") == "1" else 0
            self._side = torch.cuda.Stream(device=device, priority=prio)
        return self._side

    def _encode_image(self, images):
        eng = self.clip.engine
        if self.image_attribution is None:
            self.last_image_attribution = None
            return eng.encode_image(images)
        if self.image_attribution == "cls":
            feat, rows = eng.encode_image(images, want_cls_rows=True)
            self.last_image_attribution = (rows, None)
        else:
            feat, rows, roll = eng.encode_image(images, want_cls_rows=True, want_rollout=True)
            self.last_image_attribution = (rows, roll)
        return feat

    def _sharding(self) -> ClassSharding:
        return ClassSharding.current(self.distributed)

    def _text_gather(self, shard, n_cls, device, need_grad):
        """The fused text-feature gather (parallel.TextGather) when it applies: several ranks over NCCL on CUDA, and text features
        that are recomputed by this forward (training, or eval without the feature cache)."""
        if shard.world == 1 or not self.fused_gather:
            return None
        if self.cache_text_features and not need_grad and not self.training:
            return None                                                           # cached eval features: one NCCL all-gather per ctx version
        if self._tg is None or self._tg.n_cls != n_cls:
            if not TextGather.available(device):
                self.fused_gather = False
                return None
            try:
                self._tg = TextGather(self.clip.engine, shard, n_cls, self.clip.model.cfg.embed_dim, device)
            except Exception as e:                      # no peer access / symmetric memory on this system: every rank lands here alike
                import warnings
                warnings.warn(f"fused text-feature gather unavailable ({e!r}); using the NCCL all-gather")
                self.fused_gather, self._tg = False, None
                return None
        return self._tg

    def _adjusted_path(self):
        """'gate' / 'residual': the adjusted context is built by torch ops outside the engine (prompt_adjustor.py:38-44)."""
        return self.prompt_adjustor.method != "scale"

    def _adjusted_context(self, params):
        """model_wrapper.py:55-66 for all classes at once: attribution of the un-adjusted prompts (detached, as the hook's
        ``.detach()`` makes it in the reference), then PromptAdjustor on the [C,P,D] bank -- differentiable w.r.t. the
        context vectors and the adjustor's own parameters."""
        pl = self.prompt_learner
        bank = torch.stack(list(params), 0)                                   # [C, P, D], autograd fans the gradient back out
        if self.clip.attribution == "intended":
            attr, raw = self.clip.engine.text_attribution(pl.flat_ctx(), pl.flat_tok())
            self.clip.attention_maps[:] = [raw]
        else:
            attr = bank.new_ones(bank.shape[0], 1)                             # literal hook: attribution == 1.0
            self.clip.attention_maps[:] = [attr]
        self.last_attribution = attr
        return self.prompt_adjustor(bank, attr)            # attr [C,P] or [C,1]: broadcasts exactly like the reference's [B,1,1]

    def _text_features(self, lo, hi, need_grad, adjusted=None, gather=None):
        """Rows A6-A10 for classes [lo, hi).  With ctx unchanged and no gradient needed (evaluation: the reference
        recomputes the text side for every batch, test_cross_domain.py:84) the result is reused."""
        pl = self.prompt_learner
        if adjusted is not None:                                               # feature pass on the host-adjusted context
            return self.clip.engine.text_forward(adjusted[lo:hi].contiguous(), pl.flat_tok()[lo:hi], "literal", need_grad, **({"gather": gather} if gather else {}))
        ctx_bank = pl.flat_ctx()
        use_cache = self.cache_text_features and not need_grad and not self.training
        if use_cache:
            key = (ctx_bank.data_ptr(), tuple(p._version for p in pl.context_bank.values()), lo, hi, self.clip.attribution,
                   self.clip.engine.weight_generation)
            if self._text_cache is not None and self._text_cache[0] == key:
                return self._text_cache[1]
        out = self.clip.engine.text_forward(ctx_bank[lo:hi], pl.flat_tok()[lo:hi], self.clip.attribution, need_grad, **({"gather": gather} if gather else {}))
        self._text_cache = (key, out) if use_cache else None
        return out

    def forward(self, images, labels=None):
        pl = self.prompt_learner
        params = [pl.context_bank[k] for k in pl.context_bank.keys()]          # model_wrapper.py:47 (insertion order)
        images = images.contiguous().float()
        if images.shape[0] == 0:                                               # empty batch: cat of empty columns (:83)
            outputs = {"logits": images.new_zeros(0, pl.n_cls)}
            if labels is not None:
                nan = images.new_full((), float("nan"))                        # F.cross_entropy over an empty batch
                outputs.update({"loss": nan, "loss_cls": nan})
            return outputs
        # grad mode is off inside autograd.Function.forward, so decide here whether activations must be kept
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if self._adjusted_path():
            adj = self._adjusted_context(params)
            need_grad = torch.is_grad_enabled() and adj.requires_grad
            logits, loss = _TapClipFunction.apply(self, images, labels, need_grad, self.logit_scale, adj)
        else:
            logits, loss = _TapClipFunction.apply(self, images, labels, need_grad, self.logit_scale, *params)
        outputs = {"logits": logits}                                           # model_wrapper.py:88
        if labels is not None:                                                 # model_wrapper.py:90-93
            outputs.update({"loss": loss, "loss_cls": loss})
        if self.last_image_attribution is not None:
            rows, roll = self.last_image_attribution
            outputs["image_cls_attention"] = rows
            if roll is not None:
                outputs["image_attribution"] = roll
        return outputs
