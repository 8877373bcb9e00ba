"""Generate tests/golden/*.pt from the UNMODIFIED reference modules (build container only).

Run:  python -m oracle.make_goldens            (needs /root/reference; ~4 min on 8 cores)

For every case it
  1. imports ``models.model_wrapper.FullModel`` (and, through it, PromptLearner,
     AttributionMonitor, PromptAdjustor) straight from /root/reference — the only change
     is the hard-coded ``device='cuda'`` default of ``PromptLearner.__init__``
     (models/prompt_learner.py:7), overridden to 'cpu'; no arithmetic changes;
  2. drives it with ``StandInCLIPWrapper`` (open_clip is absent: SURVEY 8c);
  3. runs forward + ``loss.backward()`` exactly like train.py:99-104;
  4. checks that ``oracle.tapclip_oracle`` reproduces it (as-written form: bit-exact;
     de-duplicated form: to fp32 round-off) and
  5. stores the REFERENCE's outputs as the golden vectors.

Weights/inputs are regenerated from seeds at test time (same torch build in the
image), so only outputs are stored.
"""
from __future__ import annotations

import os
import sys
import time

import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.clip_standin import StandInCLIPWrapper, get_config          # noqa: E402
from oracle.tapclip_oracle import (OracleFullModel, class_names,        # noqa: E402
                                   synthetic_images, synthetic_labels)

CASES = [
    # name,            model,               B,  C,  P
    ("mini16_b4c5p4",  "mini-16",           4,  5,  4),
    ("mini16q_b3c7p5", "mini-16-quickgelu", 3,  7,  5),
    ("mini14_b2c3p16", "mini-14",           2,  3, 16),
    ("vitb16_c1",      "ViT-B-16",          8, 65, 16),     # BASELINE.json configs[0]
]
# PromptAdjustor 'gate' / 'residual' (models/prompt_adjustor.py:13-25,38-44): name, model, B, C, P, method
ADJUSTOR_CASES = [
    ("mini16_gate_b3c4p4",      "mini-16",   3, 4, 4, "gate"),
    ("minit512_resid_b2c3p5",   "mini-t512", 2, 3, 5, "residual"),       # residual_net hard-codes a 512-wide text tower
]
CTX_SEED = 4


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit("make_goldens needs /root/reference (build container only)")
    sys.path.insert(0, REF)
    from models.model_wrapper import FullModel
    from models.prompt_learner import PromptLearner
    d = list(PromptLearner.__init__.__defaults__)
    assert d[-1] == "cuda"
    d[-1] = "cpu"
    PromptLearner.__init__.__defaults__ = tuple(d)
    return FullModel


def run_reference(FullModel, model_name, B, C, P, mode, method="scale"):
    cfg = get_config(model_name)
    wrapper = StandInCLIPWrapper(model_name, device="cpu", seed=0, attribution=mode)
    torch.manual_seed(CTX_SEED)                         # PromptLearner draws ctx from the global RNG (:41)
    model = FullModel(class_names(C), wrapper, prompt_len=P, adjustor_method=method)
    model.train()                                       # train.py:91
    captured = []
    h = model.attribution_monitor.register_forward_hook(lambda m, i, o: captured.append(o.detach().clone()))
    images, labels = synthetic_images(B, cfg.image_size), synthetic_labels(B, C)
    t0 = time.time()
    out = model(images, labels)
    out["loss"].backward()
    dt = time.time() - t0
    h.remove()
    attribution = torch.stack([captured[i * B] for i in range(C)]).squeeze(1)      # sample b=0 of each class
    for i in range(C):                                  # attribution does not depend on b
        for b in range(1, B):
            assert torch.equal(captured[i * B + b], captured[i * B])
    ctx_grad = torch.stack([model.prompt_learner.context_bank[n].grad for n in class_names(C)])
    return model, {
        "logits": out["logits"].detach(), "loss": out["loss"].detach(), "attribution": attribution,
        "ctx_grad": ctx_grad, "logit_scale_grad": model.logit_scale.grad.detach().clone(),
        "adjustor_grad": {k: p.grad.detach().clone() for k, p in model.prompt_adjustor.named_parameters()},
        "adjustor_state": {k: v.detach().clone() for k, v in model.prompt_adjustor.state_dict().items()},
        "seconds_fwd_bwd": dt,
    }, (images, labels)


def check_oracle(model_name, B, C, P, mode, gold, images, labels, full_loop, method="scale"):
    wrapper = StandInCLIPWrapper(model_name, device="cpu", seed=0, attribution=mode)
    torch.manual_seed(CTX_SEED)
    orc = OracleFullModel(class_names(C), wrapper, prompt_len=P, adjustor_method=method)
    for k, v in orc.prompt_adjustor.state_dict().items():
        assert torch.equal(v, gold["adjustor_state"][k]), "adjustor initialisation differs from the reference (RNG order)"
    orc.train()
    if full_loop:
        out = orc.forward_as_written(images, labels)
        out["loss"].backward()
        g = torch.stack([orc.prompt_learner.context_bank[n].grad for n in class_names(C)])
        assert torch.equal(out["logits"], gold["logits"]), "restated loop form is not bit-exact vs the reference"
        assert torch.equal(g, gold["ctx_grad"])
        for k, p in orc.prompt_adjustor.named_parameters():
            assert torch.equal(p.grad, gold["adjustor_grad"][k]), k
        orc.zero_grad()
    out = orc.forward_dedup(images, labels, return_aux=True)
    out["loss"].backward()
    g = torch.stack([orc.prompt_learner.context_bank[n].grad for n in class_names(C)])
    err = {
        "logits": (out["logits"] - gold["logits"]).abs().max().item(),
        "loss": (out["loss"] - gold["loss"]).abs().item(),
        "attribution": (out["attribution"] - gold["attribution"]).abs().max().item(),
        "ctx_grad_rel": ((g - gold["ctx_grad"]).norm() / gold["ctx_grad"].norm()).item(),
        "logit_scale_grad": (orc.logit_scale.grad - gold["logit_scale_grad"]).abs().item(),
    }
    assert err["logits"] < 5e-5 and err["attribution"] < 1e-6 and err["ctx_grad_rel"] < 1e-4, err
    return err, out


def main():
    torch.set_num_threads(os.cpu_count())
    FullModel = import_reference()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    only_adjustor = "--adjustor-only" in sys.argv
    for name, model_name, B, C, P in ([] if only_adjustor else CASES):
        for mode in ("literal", "intended"):
            _, gold, (images, labels) = run_reference(FullModel, model_name, B, C, P, mode)
            err, dd = check_oracle(model_name, B, C, P, mode, gold, images, labels, full_loop=not name.startswith("vit"))
            gold.update({"case": name, "model_name": model_name, "B": B, "C": C, "P": P, "mode": mode,
                         "ctx_seed": CTX_SEED, "attr_raw_dedup": dd["attr_raw"].detach().clone(),
                         "text_feat_dedup": dd["text_feat"].detach().clone(), "image_feat_dedup": dd["image_feat"].detach().clone(),
                         "torch": torch.__version__, "oracle_dedup_err": err})
            path = os.path.join(ROOT, "tests", "golden", f"{name}_{mode}.pt")
            torch.save(gold, path)
            print(f"{name:16s} {mode:8s} ref fwd+bwd {gold['seconds_fwd_bwd']:.1f}s  dedup-vs-ref {err}", flush=True)
    for name, model_name, B, C, P, method in ADJUSTOR_CASES:
        for mode in ("literal", "intended"):
            _, gold, (images, labels) = run_reference(FullModel, model_name, B, C, P, mode, method)
            err, dd = check_oracle(model_name, B, C, P, mode, gold, images, labels, True, method)
            gold.update({"case": name, "model_name": model_name, "B": B, "C": C, "P": P, "mode": mode, "method": method,
                         "ctx_seed": CTX_SEED, "torch": torch.__version__, "oracle_dedup_err": err})
            torch.save(gold, os.path.join(ROOT, "tests", "golden", f"{name}_{mode}.pt"))
            print(f"{name:22s} {mode:8s} {method:8s} dedup-vs-ref {err}", flush=True)


if __name__ == "__main__":
    main()
