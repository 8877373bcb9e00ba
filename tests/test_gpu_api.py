"""Drop-in surface: constructor kwargs, state-dict keys, add_class_prompt, optimizers, errors, eval cache."""
import pytest
import torch

from helpers import build_cuda, build_oracle, ctx_grads, max_abs, rel_err
from oracle.tapclip_oracle import class_names, synthetic_images, synthetic_labels

pytestmark = pytest.mark.gpu


def _mini(mode="intended", dtype="fp32", C=5, P=4):   # fp32 parity mode: the API tests compare against the oracle tightly
    ow, om = build_oracle("mini-16", C, P, mode)
    clip, model = build_cuda("mini-16", C, P, mode, dtype, ow)
    return ow, om, clip, model


def test_state_dict_keys_and_roundtrip():
    ow, om, clip, model = _mini()
    sd = model.state_dict()
    for n in class_names(5):
        assert f"prompt_learner.context_bank.{n}" in sd
    assert "logit_scale" in sd and "clip.model.visual.conv1.weight" in sd and "clip.model.text_projection" in sd
    assert set(k for k in sd if k.startswith("clip.model.")) == set("clip.model." + k for k in ow.model.state_dict())
    names = [n for n, _ in model.named_parameters()]
    assert sum("prompt_learner.context_bank" in n for n in names) == 5          # test_cross_domain2.py:13-15 relies on this
    # legacy-key conversion path of test_cross_domain.py:46-61 -> load_state_dict(strict=False)
    new_ctx = {f"prompt_learner.context_bank.{n}": torch.full((4, 256), float(i)) for i, n in enumerate(class_names(5))}
    model.load_state_dict(new_ctx, strict=False)
    flat = model.prompt_learner.flat_ctx()
    assert torch.equal(flat[3], torch.full((4, 256), 3.0, device="cuda"))


def test_add_class_prompt_after_construction_matches_fresh_model():
    ow, om, clip, model = _mini(C=3)
    torch.manual_seed(99)
    for n in class_names(12)[3:]:
        model.prompt_learner.add_class_prompt(n)              # test_cross_domain.py:65-67; grows the flat bank
    model.prompt_learner.add_class_prompt(class_names(12)[0])  # existing -> no-op
    assert model.prompt_learner.n_cls == 12
    images = synthetic_images(2, 64).cuda()
    model.eval()
    with torch.no_grad():
        l12 = model(images)["logits"]
    assert l12.shape == (2, 12)
    # oracle with the same 12 ctx vectors
    from oracle.tapclip_oracle import OracleFullModel
    om2 = OracleFullModel(class_names(12), ow, prompt_len=4)
    with torch.no_grad():
        for n in class_names(12):
            om2.prompt_learner.context_bank[n].copy_(model.prompt_learner.context_bank[n].cpu())
        ref = om2.forward_dedup(images.cpu())["logits"]
    assert max_abs(l12, ref) < 1e-4


def test_prompt_learner_forward_shape_and_errors():
    import tapclip_b200 as tb
    ow, om, clip, model = _mini()
    assert model.prompt_learner().shape == (5, 4 + 77, 256)
    assert max_abs(model.prompt_learner(), om.prompt_learner()) == 0.0
    with pytest.raises(ValueError):
        tb.PromptAdjustor(method="bogus")                      # prompt_adjustor.py:47
    with pytest.raises(ValueError):
        tb.CLIPWrapper("ViT-Z-99", None, "cuda")
    with pytest.raises(ValueError):
        model(torch.zeros(2, 3, 32, 32, device="cuda"))        # wrong image size
    with pytest.raises((ValueError, RuntimeError)):
        clip.engine.load_state_dict({"visual.conv1.weight": torch.zeros(3, 3, device="cuda")})


def test_attribution_monitor_module():
    import tapclip_b200 as tb
    g = torch.Generator(device="cuda").manual_seed(0)
    attn = torch.softmax(torch.randn(6, 20, 20, device="cuda", generator=g), -1)
    ref = torch.softmax(attn[:, :5, 19], dim=-1)
    assert max_abs(tb.AttributionMonitor(5)(attn), ref) < 1e-6
    assert max_abs(tb.AttributionMonitor(5, normalize=False)(attn), attn[:, :5, 19]) < 1e-7


@pytest.mark.parametrize("fused", [False, True])
def test_three_train_steps_track_the_oracle(fused):
    """train.py:65-67,99-105: AdamW on prompt_learner.parameters(); losses and ctx follow the CPU oracle."""
    import tapclip_b200 as tb
    ow, om, clip, model = _mini(mode="intended", dtype="fp32")
    images, labels = synthetic_images(4, 64), synthetic_labels(4, 5)
    opt_ref = torch.optim.AdamW(om.prompt_learner.parameters(), lr=2e-3, weight_decay=0.01)
    opt = tb.FusedAdamW(model, lr=2e-3, weight_decay=0.01) if fused else \
        torch.optim.AdamW(model.prompt_learner.parameters(), lr=2e-3, weight_decay=0.01)
    model.train(); om.train()
    for step in range(3):
        lr = om.forward_dedup(images, labels)["loss"]
        opt_ref.zero_grad(); lr.backward(); opt_ref.step()
        lc = model(images.cuda(), labels.cuda())["loss"]
        opt.zero_grad(); lc.backward(); opt.step()
        assert abs(lc.item() - lr.item()) < 2e-4, (step, lc.item(), lr.item())
    ref_ctx = torch.stack([om.prompt_learner.context_bank[n].detach() for n in class_names(5)])
    assert max_abs(model.prompt_learner.flat_ctx(), ref_ctx) < 2e-4


def test_eval_text_feature_cache_and_argmax():
    ow, om, clip, model = _mini(mode="intended", dtype="fp32")
    images, labels = synthetic_images(4, 64).cuda(), synthetic_labels(4, 5).cuda()
    model.eval()
    with torch.no_grad():
        n0 = clip.engine.launch_count
        l1 = model(images)["logits"]
        n1 = clip.engine.launch_count
        l2 = model(images)["logits"]
        n2 = clip.engine.launch_count
    assert torch.equal(l1, l2) and (n2 - n1) < (n1 - n0)         # text side reused when ctx is unchanged
    with torch.no_grad():
        model.prompt_learner.context_bank[class_names(5)[0]].add_(1.0)
        l3 = model(images)["logits"]
    assert not torch.equal(l3, l1)                               # in-place ctx update invalidates the cache
    pred, correct = clip.engine.argmax_count(l1, labels)
    assert torch.equal(pred, l1.argmax(1)) and correct.item() == (l1.argmax(1) == labels).sum().item()


def test_evaluate_accuracy_with_device_side_per_class_counters(capsys):
    """utils/eval_metrics.py:7-41,44-73 as a drop-in: same return values and report as the reference's host loop, with the
    argmax and the overall / per-class counters kept on the device (one D2H per epoch)."""
    import tapclip_b200 as tb
    ow, om, clip, model = _mini(mode="intended", dtype="fp32")
    C = 5
    g = torch.Generator().manual_seed(3)
    loader = [(torch.randn(n, 3, 64, 64, generator=g), torch.randint(0, C, (n,), generator=g)) for n in (7, 4, 1, 9)]
    # the reference's loop (eval_metrics.py:14-29), restated on the host from this model's own logits
    correct = total = 0
    cls_ok, cls_n = [0] * C, [0] * C
    model.eval()
    with torch.no_grad():
        for images, labels in loader:
            preds = model(images.cuda())["logits"].argmax(1).cpu()
            correct += int((preds == labels).sum()); total += labels.numel()
            for t, p in zip(labels.tolist(), preds.tolist()):
                cls_n[t] += 1; cls_ok[t] += int(t == p)
    acc = tb.evaluate_accuracy(model, loader, "cuda")
    report = capsys.readouterr().out
    assert abs(acc - 100.0 * correct / total) < 1e-9
    for c in range(C):
        if cls_n[c]:
            assert f" - Class {c:2d}: {100.0 * cls_ok[c] / cls_n[c]:.2f}% ({cls_ok[c]}/{cls_n[c]})" in report
    per = tb.evaluate_per_class_accuracy(model, loader, "cuda", class_names=class_names(C))
    assert per == {class_names(C)[c]: 100.0 * cls_ok[c] / cls_n[c] for c in range(C) if cls_n[c]}
    assert tb.evaluate_accuracy(model, [], "cuda") == 0.0                      # empty loader (eval_metrics.py:31)
    # counters accumulate across calls and ignore out-of-range labels for the per-class arrays
    cnt = (torch.zeros(1, dtype=torch.int32, device="cuda"), torch.zeros(C, dtype=torch.int32, device="cuda"),
           torch.zeros(C, dtype=torch.int32, device="cuda"))
    logits = torch.randn(6, C, device="cuda")
    logits[2, 3] = float("nan")                                               # torch.argmax: NaN is the maximum
    lab = torch.tensor([0, 1, 3, 4, 2, 0], device="cuda")
    for _ in range(2):
        pred, _ = clip.engine.argmax_count(logits, lab, counters=cnt)
    assert torch.equal(pred, logits.argmax(1))
    assert cnt[0].item() == 2 * int((pred == lab).sum()) and cnt[2].sum().item() == 12


def test_stale_backward_is_refused():
    """One CLIPWrapper shared by two FullModels (train + EMA/teacher): the engine holds ONE set of saved text activations, so a
    backward that follows another model's forward must fail loudly instead of differentiating the wrong forward."""
    import tapclip_b200 as tb
    ow, om, clip, model = _mini(mode="intended", dtype="fp32")
    torch.manual_seed(5)
    other = tb.FullModel(class_names(5), clip, prompt_len=model.prompt_len)
    images, labels = synthetic_images(4, 64).cuda(), synthetic_labels(4, 5).cuda()
    model.train(); other.train()
    out_a = model(images, labels)
    out_b = other(images, labels)                       # overwrites the saved activations of out_a's forward
    with pytest.raises(Exception, match="stale backward"):
        out_a["loss"].backward()
    out_b["loss"].backward()                            # the latest forward is still differentiable
    assert all(p.grad is not None for p in other.prompt_learner.context_bank.values())


@pytest.mark.parametrize("dtype", ["fp32", "mixed"])
def test_standard_encode_text_path(dtype):
    """CLIPWrapper.encode_text (clip_wrapper.py:49-51): positional embedding + causal mask + ln_final + EOT pooling."""
    ow, _ = build_oracle("mini-16", 2, 4, "literal")
    clip, _ = build_cuda("mini-16", 2, 4, "literal", dtype, ow)
    texts = ["a photo of a cat", "a photo of a very tall giraffe eating leaves", "x"]
    ids = ow.get_tokenizer()(texts)
    with torch.no_grad():
        ref = ow.encode_text(ids)
    out = clip.encode_text(ids.cuda())
    assert clip.attention_maps == []                       # encode_text resets the probe cache (clip_wrapper.py:50)
    tol = 1e-4 if dtype == "fp32" else 2e-2
    assert max_abs(out, ref) < tol * max(1.0, ref.abs().max().item())


def test_logits_only_backward_path():
    """A caller that builds its own loss from outputs['logits'] (not outputs['loss']) still gets ctx gradients."""
    ow, om, clip, model = _mini(mode="literal", dtype="fp32")
    images, labels = synthetic_images(4, 64), synthetic_labels(4, 5)
    model.train(); om.train()
    (model(images.cuda())["logits"].logsumexp(1).sum()).backward()
    (om.forward_dedup(images)["logits"].logsumexp(1).sum()).backward()
    ref = torch.stack([om.prompt_learner.context_bank[n].grad for n in class_names(5)])
    assert ((ctx_grads(model, 5) - ref).norm() / ref.norm()).item() < 2e-3


@pytest.mark.parametrize("kind", ["cls", "rollout"])
def test_forward_with_image_attribution(kind):
    """FullModel(image_attribution=...) (north-star extension, SURVEY 8a row A-ext): the image pass that feeds the logits
    also emits the per-layer CLS-row attention (and the rollout map); logits are unchanged by the probes."""
    import tapclip_b200 as tb
    from oracle.clip_standin import vision_attention_rollout, vision_cls_attention
    ow, om, clip, plain = _mini(mode="literal", dtype="fp32")
    torch.manual_seed(4)
    model = tb.FullModel(class_names(5), clip, prompt_len=4, image_attribution=kind)
    images = synthetic_images(4, 64)
    with torch.no_grad():
        out = model(images.cuda())
        ref_logits = plain(images.cuda())["logits"]
        _, rows_ref = vision_cls_attention(ow.model, images)                    # [B, L, H, N]
    assert torch.equal(out["logits"], ref_logits)
    assert out["image_cls_attention"].shape == rows_ref.shape
    assert max_abs(out["image_cls_attention"], rows_ref) < 1e-4
    if kind == "rollout":
        with torch.no_grad():
            roll_ref = vision_attention_rollout(ow.model, images)
        assert max_abs(out["image_attribution"], roll_ref) < 1e-4
    else:
        assert "image_attribution" not in out
    with pytest.raises(ValueError):
        tb.FullModel(class_names(5), clip, prompt_len=4, image_attribution="gradcam")


@pytest.mark.parametrize("dtype", ["fp32", "mixed"])
def test_fused_head_matches_the_separate_launches(dtype, monkeypatch):
    """K4 (head.cu): pool + text_projection + L2-norm in one launch, image L2-norm + logits + cross-entropy + batch mean in one
    launch (last-CTA reduction: deterministic), and their two backward kernels -- against the 7 + 5 separate launches
    (TAPCLIP_FUSE_HEAD=0) and the oracle; a label outside [0, C) yields NaN instead of an out-of-bounds read."""
    res = {}
    for fuse in ("1", "0"):
        monkeypatch.setenv("TAPCLIP_FUSE_HEAD", fuse)
        ow, om, clip, model = _mini(mode="intended", dtype=dtype, C=7, P=5)
        images, labels = synthetic_images(6, 64).cuda(), synthetic_labels(6, 7).cuda()
        model.train()
        n0 = clip.engine.launch_count
        out = model(images, labels)
        out["loss"].backward()
        torch.cuda.synchronize()
        res[fuse] = (out["logits"].detach().cpu(), out["loss"].item(), ctx_grads(model, 7), model.logit_scale.grad.item(),
                     clip.engine.launch_count - n0)
        out2 = model(images, labels)                               # deterministic: bit-identical loss on a second run
        assert out2["loss"].item() == out["loss"].item()
        if fuse == "1":
            om.train()
            ref = om.forward_dedup(images.cpu(), labels.cpu())
            assert abs(out["loss"].item() - ref["loss"].item()) < (1e-5 if dtype == "fp32" else 1e-2)
            bad = labels.clone(); bad[2] = 7
            assert torch.isnan(model(images, bad)["loss"])
            bad[2] = -100
            assert torch.isnan(model(images, bad)["loss"])
    tol = 1e-5 if dtype == "fp32" else 5e-3
    assert max_abs(res["1"][0], res["0"][0]) < tol and abs(res["1"][1] - res["0"][1]) < tol
    assert rel_err(res["1"][2], res["0"][2]) < (1e-5 if dtype == "fp32" else 2e-2)
    assert abs(res["1"][3] - res["0"][3]) < (1e-5 if dtype == "fp32" else 5e-4)   # sum of cancelling terms: follows the logit differences
    assert res["1"][4] <= res["0"][4] - 8                            # 7 -> 2 launches forward, 5 -> 2 backward
