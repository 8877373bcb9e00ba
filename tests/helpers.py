"""Shared helpers for the GPU parity tests (tests only — the oracle is the checker, never the product path)."""
import ctypes as C
import os

import torch

from oracle.clip_standin import StandInCLIPWrapper, get_config
from oracle.tapclip_oracle import OracleFullModel, class_names, synthetic_images, synthetic_labels

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CTX_SEED = 4


def load_golden(case, mode):
    return torch.load(os.path.join(GOLDEN_DIR, f"{case}_{mode}.pt"), map_location="cpu", weights_only=False)


def build_oracle(model_name, C_, P, mode, seed=0, method="scale"):
    wrapper = StandInCLIPWrapper(model_name, device="cpu", seed=seed, attribution=mode)
    torch.manual_seed(CTX_SEED)
    model = OracleFullModel(class_names(C_), wrapper, prompt_len=P, adjustor_method=method)
    return wrapper, model


def build_cuda(model_name, C_, P, mode, dtype, oracle_wrapper, method="scale"):
    """tapclip_b200 model with the oracle's weights and the same ctx (+ adjustor) draw (global CPU RNG, seed 4)."""
    import tapclip_b200 as tb
    clip = tb.CLIPWrapper(model_name, None, "cuda", state_dict=oracle_wrapper.model.state_dict(), attribution=mode, dtype=dtype,
                          tokenizer="synthetic")      # the oracle's weights are random-init: same hashed ids on both sides
    torch.manual_seed(CTX_SEED)
    model = tb.FullModel(class_names(C_), clip, prompt_len=P, adjustor_method=method)
    return clip, model


def ctx_grads(model, C_):
    return torch.stack([model.prompt_learner.context_bank[n].grad.detach().cpu() for n in class_names(C_)])


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def max_abs(a, b):
    return (a.detach().float().cpu() - b.detach().float().cpu()).abs().max().item()


def top1_agreement(logits, ref_logits, margin):
    """(raw agreement, agreement restricted to samples whose reference top1-top2 gap exceeds `margin`, #such samples)."""
    logits, ref_logits = logits.detach().float().cpu(), ref_logits.detach().float().cpu()
    agree = logits.argmax(1) == ref_logits.argmax(1)
    top2 = ref_logits.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > margin
    raw = agree.float().mean().item()
    filt = agree[clear].float().mean().item() if clear.any() else 1.0
    return raw, filt, int(clear.sum())
