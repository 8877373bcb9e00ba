"""The CPU oracle against the golden vectors produced by the UNMODIFIED reference modules (oracle/make_goldens.py).

These pin the restatement in oracle/tapclip_oracle.py to the reference's own FullModel / PromptLearner /
AttributionMonitor / PromptAdjustor.  (The open_clip model underneath is restated — see test_oracle_hf.py.)
"""
import os

import pytest
import torch

from helpers import build_oracle, load_golden
from oracle.clip_standin import get_config
from oracle.tapclip_oracle import class_names, synthetic_images, synthetic_labels

MINI = ["mini16_b4c5p4", "mini16q_b3c7p5", "mini14_b2c3p16"]


def _grads(m, C):
    return torch.stack([m.prompt_learner.context_bank[n].grad for n in class_names(C)])


@pytest.mark.parametrize("mode", ["literal", "intended"])
@pytest.mark.parametrize("case", MINI)
def test_loop_form_is_bit_exact(case, mode):
    gold = load_golden(case, mode)
    B, C, P = gold["B"], gold["C"], gold["P"]
    _, m = build_oracle(gold["model_name"], C, P, mode)
    m.train()
    images, labels = synthetic_images(B, get_config(gold["model_name"]).image_size), synthetic_labels(B, C)
    out = m.forward_as_written(images, labels)
    out["loss"].backward()
    assert torch.equal(out["logits"], gold["logits"])
    assert torch.equal(out["loss"], gold["loss"])
    assert torch.equal(_grads(m, C), gold["ctx_grad"])
    assert torch.equal(m.logit_scale.grad, gold["logit_scale_grad"])


@pytest.mark.parametrize("mode", ["literal", "intended"])
@pytest.mark.parametrize("case", MINI + ["vitb16_c1"])
def test_dedup_form_matches_reference(case, mode):
    gold = load_golden(case, mode)
    B, C, P = gold["B"], gold["C"], gold["P"]
    torch.set_num_threads(os.cpu_count())
    _, m = build_oracle(gold["model_name"], C, P, mode)
    m.train()
    images, labels = synthetic_images(B, get_config(gold["model_name"]).image_size), synthetic_labels(B, C)
    out = m.forward_dedup(images, labels, return_aux=True)
    out["loss"].backward()
    assert (out["logits"] - gold["logits"]).abs().max().item() < 5e-5
    assert abs(out["loss"].item() - gold["loss"].item()) < 1e-5
    assert (out["attribution"] - gold["attribution"]).abs().max().item() < 1e-6
    g = _grads(m, C)
    assert ((g - gold["ctx_grad"]).norm() / gold["ctx_grad"].norm()).item() < 1e-4
    assert abs(m.logit_scale.grad.item() - gold["logit_scale_grad"].item()) < 1e-5
    if mode == "literal":
        assert torch.equal(out["attribution"], torch.ones(C, 1))          # attribution == 1.0 exactly
    else:
        assert out["attribution"].shape == (C, P)
        assert (out["attribution"].sum(-1) - 1).abs().max().item() < 1e-6


def test_text_side_does_not_depend_on_the_sample():
    """rows of the batch-B text pass are identical, which is what de-duplication relies on."""
    ow, m = build_oracle("mini-16", 3, 4, "intended")
    raw_prompt = m.prompt_learner()
    x = raw_prompt[1].unsqueeze(0).expand(4, -1, -1)
    y = ow.model.transformer(x)
    assert torch.equal(y[0], y[3])


ADJUSTOR = ["mini16_gate_b3c4p4", "minit512_resid_b2c3p5"]


@pytest.mark.parametrize("mode", ["literal", "intended"])
@pytest.mark.parametrize("case", ADJUSTOR)
def test_gate_and_residual_adjustors_are_bit_exact(case, mode):
    """PromptAdjustor 'gate' / 'residual' (models/prompt_adjustor.py:13-25,38-44) driven through the reference's FullModel:
    logits, ctx gradients and the adjustor networks' own gradients."""
    gold = load_golden(case, mode)
    B, C, P = gold["B"], gold["C"], gold["P"]
    _, m = build_oracle(gold["model_name"], C, P, mode, method=gold["method"])
    for k, v in m.prompt_adjustor.state_dict().items():
        assert torch.equal(v, gold["adjustor_state"][k])                  # same RNG order as the reference's constructor
    m.train()
    images, labels = synthetic_images(B, get_config(gold["model_name"]).image_size), synthetic_labels(B, C)
    out = m.forward_as_written(images, labels)
    out["loss"].backward()
    assert torch.equal(out["logits"], gold["logits"]) and torch.equal(_grads(m, C), gold["ctx_grad"])
    for k, p in m.prompt_adjustor.named_parameters():
        assert torch.equal(p.grad, gold["adjustor_grad"][k]), k
    m.zero_grad()
    out = m.forward_dedup(images, labels)
    out["loss"].backward()
    assert (out["logits"] - gold["logits"]).abs().max().item() < 5e-5
    assert ((_grads(m, C) - gold["ctx_grad"]).norm() / gold["ctx_grad"].norm()).item() < 1e-4
    for k, p in m.prompt_adjustor.named_parameters():
        ref = gold["adjustor_grad"][k]
        assert ((p.grad - ref).norm() / ref.norm().clamp_min(1e-12)).item() < 1e-4, k
