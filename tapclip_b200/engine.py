"""Thin Python handle around libtapclip's engine: tensors in, tensors out, all work on the current stream."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .configs import ModelConfig


def _check_cuda_f32(t: torch.Tensor, name: str):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise ValueError(f"{name} must be a contiguous fp32 CUDA tensor (got {t.dtype}, {t.device}, contiguous={t.is_contiguous()})")


class Engine:
    """One engine per device.  ``dtype``: 'mixed' | 'bf16' (tcgen05 tensor-core paths) or 'fp32' (SIMT parity mode)."""

    def __init__(self, cfg: ModelConfig, dtype: str = "mixed", device="cuda"):
        if dtype not in ("fp32", "bf16", "mixed"):
            raise ValueError(f"dtype must be 'mixed', 'bf16' or 'fp32', got {dtype!r}")
        if not torch.cuda.is_available():
            raise _lib.TapclipError("tapclip_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg, self.dtype = cfg, dtype
        self.device = torch.device(device if device != "cuda" else f"cuda:{torch.cuda.current_device()}")
        c = _lib.TapclipConfig(cfg.image_size, cfg.patch_size, cfg.vision_width, cfg.vision_layers, cfg.vision_heads,
                               cfg.text_width, cfg.text_layers, cfg.text_heads, cfg.embed_dim, cfg.context_length,
                               _lib.ACT["quick_gelu" if cfg.quick_gelu else "gelu_erf"], _lib.DTYPE[dtype])
        self._h = C.c_void_p()
        self.last_forward_token = 0
        self.weight_generation = 0                       # bumped by load_state_dict: part of FullModel's text-feature cache key
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tapclip_create(C.byref(c), C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self.lib.tapclip_destroy(h)
            except Exception:
                pass
            self._h = C.c_void_p()

    # ---- weights -----------------------------------------------------------------------------------
    def load_state_dict(self, state_dict):
        """One tapclip_load_weight per entry (open_clip key names); replaces clip_wrapper.py:14-15."""
        with torch.cuda.device(self.device):
            for name, t in state_dict.items():
                w = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
                shape = (C.c_int64 * max(w.dim(), 1))(*w.shape)
                _lib.check(self.lib.tapclip_load_weight(self._h, name.encode(), _lib.ptr(w), w.dim(), shape, _lib.stream_ptr()))
            torch.cuda.current_stream().synchronize()      # staging copies `w` may be freed after this; folded operands are built
            _lib.check(self.lib.tapclip_weights_complete(self._h))
        self.weight_generation += 1

    # ---- hot path -------------------------------------------------------------------------------------
    def encode_image(self, images: torch.Tensor, want_cls_rows: bool = False, want_rollout: bool = False):
        _check_cuda_f32(images, "images")
        cfg = self.cfg
        if images.dim() != 4 or images.shape[1] != 3 or images.shape[2] != cfg.image_size or images.shape[3] != cfg.image_size:
            raise ValueError(f"images must be [B,3,{cfg.image_size},{cfg.image_size}], got {tuple(images.shape)}")
        B = images.shape[0]
        feat = torch.empty(B, cfg.embed_dim, device=images.device, dtype=torch.float32)
        rows = None
        if want_cls_rows:
            rows = torch.empty(B, cfg.vision_layers, cfg.vision_heads, cfg.vision_tokens, device=images.device, dtype=torch.float32)
        roll = torch.empty(B, cfg.vision_tokens - 1, device=images.device, dtype=torch.float32) if want_rollout else None
        _lib.check(self.lib.tapclip_encode_image(self._h, _lib.ptr(images), B, _lib.ptr(feat), _lib.ptr(rows), _lib.ptr(roll),
                                                 _lib.stream_ptr()))
        if want_rollout:
            return (feat, rows, roll) if want_cls_rows else (feat, roll)
        return (feat, rows) if want_cls_rows else feat

    def encode_text(self, token_ids: torch.Tensor):
        if token_ids.dtype != torch.int64 or not token_ids.is_cuda or token_ids.dim() != 2 or token_ids.shape[1] != self.cfg.context_length:
            raise ValueError(f"token ids must be an int64 CUDA tensor [S, {self.cfg.context_length}], got {tuple(token_ids.shape)} {token_ids.dtype}")
        ids = token_ids.contiguous()
        feat = torch.empty(ids.shape[0], self.cfg.embed_dim, device=ids.device, dtype=torch.float32)
        _lib.check(self.lib.tapclip_encode_text(self._h, _lib.ptr(ids), ids.shape[0], _lib.ptr(feat), _lib.stream_ptr()))
        return feat

    def set_text_gather(self, peer_ptrs, rank: int, n_cls_total: int):
        """K5: symmetric buffers of all ranks (device pointers as mapped into this process) for the fused text-feature gather."""
        world = len(peer_ptrs)
        arr = (C.c_void_p * max(world, 1))(*[C.c_void_p(int(p)) for p in peer_ptrs])
        _lib.check(self.lib.tapclip_text_gather_config(self._h, arr, world, rank, n_cls_total))

    def text_forward(self, ctx: torch.Tensor, tok: torch.Tensor, mode: str, save_for_backward: bool, gather=None):
        """``gather`` = (row_lo, epoch): also store the features into every rank's symmetric buffer (see set_text_gather)."""
        _check_cuda_f32(ctx, "ctx")
        _check_cuda_f32(tok, "tok")
        cfg = self.cfg
        Cn, P, D = ctx.shape
        if tok.shape != (Cn, cfg.context_length, D) or D != cfg.text_width:
            raise ValueError(f"Unexpected token shape: {tuple(tok.shape)} for ctx {tuple(ctx.shape)}")
        pa = P if mode == "intended" else 1
        attr = torch.empty(Cn, pa, device=ctx.device, dtype=torch.float32)
        raw = torch.empty(Cn, pa, device=ctx.device, dtype=torch.float32) if mode == "intended" else None
        feat = torch.empty(Cn, cfg.embed_dim, device=ctx.device, dtype=torch.float32)
        token = C.c_int64(0)
        _lib.check(self.lib.tapclip_text_forward(self._h, _lib.ptr(ctx), _lib.ptr(tok), Cn, P, _lib.ATTR_MODE[mode],
                                                 1 if save_for_backward else 0, _lib.ptr(raw), _lib.ptr(attr), _lib.ptr(feat),
                                                 C.byref(token), gather[0] if gather else 0, gather[1] if gather else 0, _lib.stream_ptr()))
        self.last_forward_token = int(token.value)      # identifies the saved activations; text_backward(token=...) checks it
        return feat, attr, raw

    def text_attribution(self, ctx: torch.Tensor, tok: torch.Tensor):
        """The INTENDED attribution pass alone (rows A7/A8): returns (attr [C,P], raw [C,P])."""
        _check_cuda_f32(ctx, "ctx")
        _check_cuda_f32(tok, "tok")
        Cn, P, D = ctx.shape
        if tok.shape != (Cn, self.cfg.context_length, D) or D != self.cfg.text_width:
            raise ValueError(f"Unexpected token shape: {tuple(tok.shape)} for ctx {tuple(ctx.shape)}")
        attr = torch.empty(Cn, P, device=ctx.device, dtype=torch.float32)
        raw = torch.empty(Cn, P, device=ctx.device, dtype=torch.float32)
        _lib.check(self.lib.tapclip_text_forward(self._h, _lib.ptr(ctx), _lib.ptr(tok), Cn, P, _lib.ATTR_MODE["attribution_only"], 0,
                                                 _lib.ptr(raw), _lib.ptr(attr), None, None, 0, 0, _lib.stream_ptr()))
        return attr, raw

    def logits(self, img_feat, text_feat, logit_scale, labels=None, inv_batch_total=None, gather_epoch=0):
        B, Cn = img_feat.shape[0], text_feat.shape[0]
        dev = img_feat.device
        img_norm = torch.empty_like(img_feat)
        logits = torch.empty(B, Cn, device=dev, dtype=torch.float32)
        loss = dlogits = None
        if labels is not None:
            if labels.dtype != torch.int64 or not labels.is_cuda:
                raise ValueError("labels must be an int64 CUDA tensor")
            loss = torch.zeros((), device=dev, dtype=torch.float32)
            dlogits = torch.empty_like(logits)
            labels = labels.contiguous()
        inv = float(inv_batch_total) if inv_batch_total is not None else (1.0 / max(B, 1))
        _lib.check(self.lib.tapclip_logits(self._h, _lib.ptr(img_feat), _lib.ptr(text_feat), _lib.ptr(logit_scale), _lib.ptr(labels),
                                           B, Cn, inv, _lib.ptr(img_norm), _lib.ptr(logits), _lib.ptr(loss), _lib.ptr(dlogits),
                                           int(gather_epoch), _lib.stream_ptr()))
        return logits, loss, dlogits, img_norm

    def logits_backward(self, dlogits, logits, img_norm, logit_scale):
        B, Cn = logits.shape
        d_text = torch.empty(Cn, self.cfg.embed_dim, device=logits.device, dtype=torch.float32)
        d_scale = torch.zeros((), device=logits.device, dtype=torch.float32)
        _lib.check(self.lib.tapclip_logits_backward(self._h, _lib.ptr(dlogits), _lib.ptr(logits), _lib.ptr(img_norm),
                                                    _lib.ptr(logit_scale), B, Cn, _lib.ptr(d_text), _lib.ptr(d_scale),
                                                    _lib.stream_ptr()))
        return d_text, d_scale

    def text_backward(self, d_text_feat, n_cls: int, prompt_len: int, token: int = 0):
        """``token``: ``last_forward_token`` of the forward being differentiated; a forward that ran on this engine since then
        has replaced the saved activations, and the call raises instead of using them (0 = no check)."""
        _check_cuda_f32(d_text_feat, "d_text_feat")
        dctx = torch.empty(n_cls, prompt_len, self.cfg.text_width, device=d_text_feat.device, dtype=torch.float32)
        _lib.check(self.lib.tapclip_text_backward(self._h, _lib.ptr(d_text_feat), _lib.ptr(dctx), int(token), n_cls, prompt_len,
                                                  _lib.stream_ptr()))
        return dctx

    def adamw_step(self, param, grad, exp_avg, exp_avg_sq, lr, betas, eps, weight_decay, step):
        for t, n in ((param, "param"), (grad, "grad"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
            _check_cuda_f32(t, n)
        _lib.check(self.lib.tapclip_adamw_step(self._h, _lib.ptr(param), _lib.ptr(grad), _lib.ptr(exp_avg), _lib.ptr(exp_avg_sq),
                                               param.numel(), lr, betas[0], betas[1], eps, weight_decay, step, _lib.stream_ptr()))

    def argmax_count(self, logits, labels=None, counters=None):
        """argmax over classes (+ accuracy counters).  ``counters``: optional ``(correct [1], class_correct [C], class_total [C])``
        int32 CUDA tensors that are ACCUMULATED into across calls (eval_metrics.evaluate_accuracy keeps one set per epoch)."""
        B, Cn = logits.shape
        pred = torch.empty(B, device=logits.device, dtype=torch.int64)
        correct = cls_ok = cls_n = None
        if labels is not None:
            if labels.dtype != torch.int64 or not labels.is_cuda:
                raise ValueError("labels must be an int64 CUDA tensor")
            labels = labels.contiguous()
            if counters is not None:
                correct, cls_ok, cls_n = counters
                for t in counters:
                    if t.dtype != torch.int32 or not t.is_cuda or not t.is_contiguous():
                        raise ValueError("accuracy counters must be contiguous int32 CUDA tensors")
                if cls_ok.numel() != Cn or cls_n.numel() != Cn:
                    raise ValueError(f"per-class counters must have {Cn} entries")
            else:
                correct = torch.zeros((), device=logits.device, dtype=torch.int32)
        _lib.check(self.lib.tapclip_argmax_count(self._h, _lib.ptr(logits.contiguous()), _lib.ptr(labels), B, Cn, _lib.ptr(pred),
                                                 _lib.ptr(correct), _lib.ptr(cls_ok), _lib.ptr(cls_n), _lib.stream_ptr()))
        return pred, correct

    def profile(self, enable: bool):
        _lib.check(self.lib.tapclip_profile(self._h, 1 if enable else 0))

    def profile_report(self) -> dict:
        import json
        return json.loads(self.lib.tapclip_profile_report(self._h).decode())

    @property
    def launch_count(self) -> int:
        return int(self.lib.tapclip_launch_count(self._h))

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.tapclip_workspace_bytes(self._h))
