"""Host-side logic of the drop-in on CPU (engine replaced by tests/fake_engine.FakeEngine): configs, tokenizer,
flat ctx bank, autograd wiring, state-dict surface — compared against the oracle / golden vectors."""
import pytest
import torch

from fake_engine import FakeWrapper
from helpers import build_oracle, load_golden
from oracle.clip_standin import SyntheticTokenizer as OracleTok, get_config
from oracle.tapclip_oracle import class_names, synthetic_images, synthetic_labels


def test_configs_and_flops_match_the_survey():
    from tapclip_b200.configs import flops_per_image, flops_per_text_sequence, get_model_config
    for name in ("ViT-B-32", "ViT-B-16", "ViT-L-14-336", "ViT-B-16-quickgelu", "mini-16", "mini-14", "mini-t512", "mini-n197"):
        a, b = get_model_config(name), get_config(name)
        for f in ("embed_dim", "image_size", "patch_size", "vision_width", "vision_layers", "vision_heads", "text_width",
                  "text_layers", "text_heads", "context_length", "vocab_size", "quick_gelu"):
            assert getattr(a, f) == getattr(b, f), (name, f)
    with pytest.raises(ValueError):
        get_model_config("ViT-Z-99")
    b16 = get_model_config("ViT-B-16")
    assert abs(flops_per_image(b16) / 1e9 - 35.127) < 0.01
    assert abs(flops_per_image(get_model_config("ViT-L-14-336")) / 1e9 - 381.92) < 0.05
    assert abs(flops_per_text_sequence(b16, 93) / 1e9 - 7.234) < 0.002
    assert abs(flops_per_text_sequence(b16, 82) / 1e9 - 6.357) < 0.002


def test_tokenizer_contract():
    from tapclip_b200.tokenizer import SyntheticTokenizer
    a, b = SyntheticTokenizer(77), OracleTok(77)
    for text in ("a photo of a class_007", "a photo of a Alarm Clock", "a photo of a " + "very " * 90 + "long"):
        ta, tb_ = a(text), b(text)
        assert ta.shape == (1, 77) and ta.dtype == torch.long and torch.equal(ta, tb_)
        assert ta[0, 0] == 49406 and (ta == 49407).sum() == 1
    assert a(["x", "y z"]).shape == (2, 77)


def test_tokenizer_choice_follows_the_weights():
    """Real weights need the vocabulary they were trained with (the reference always uses open_clip.get_tokenizer,
    clip_wrapper.py:27): without open_clip and without an explicit tokenizer the wrapper refuses instead of hashing words."""
    from tapclip_b200.clip_wrapper import CLIPWrapper
    from tapclip_b200.configs import get_model_config
    from tapclip_b200.tokenizer import SyntheticTokenizer
    cfg = get_model_config("mini-16")
    assert isinstance(CLIPWrapper._pick_tokenizer(None, "mini-16", cfg, real_weights=False), SyntheticTokenizer)
    assert isinstance(CLIPWrapper._pick_tokenizer("synthetic", "mini-16", cfg, real_weights=True), SyntheticTokenizer)
    mine = lambda s: torch.zeros(1, 77, dtype=torch.long)
    assert CLIPWrapper._pick_tokenizer(mine, "mini-16", cfg, real_weights=True) is mine
    try:
        import open_clip  # noqa: F401
    except ImportError:
        with pytest.raises(RuntimeError, match="open_clip"):
            CLIPWrapper._pick_tokenizer(None, "ViT-B-16", cfg, real_weights=True)


def _models(case, mode):
    import tapclip_b200 as tb
    gold = load_golden(case, mode)
    ow, om = build_oracle(gold["model_name"], gold["C"], gold["P"], mode)
    torch.manual_seed(4)
    model = tb.FullModel(class_names(gold["C"]), FakeWrapper(ow, mode), prompt_len=gold["P"])
    return gold, ow, om, model


@pytest.mark.parametrize("mode", ["literal", "intended"])
@pytest.mark.parametrize("case", ["mini16_b4c5p4", "mini14_b2c3p16"])
def test_fullmodel_host_path_matches_reference_golden(case, mode):
    gold, ow, om, model = _models(case, mode)
    B, C = gold["B"], gold["C"]
    images, labels = synthetic_images(B, get_config(gold["model_name"]).image_size), synthetic_labels(B, C)
    model.train()
    out = model(images, labels)
    assert set(out) == {"logits", "loss", "loss_cls"}
    out["loss"].backward()
    g = torch.stack([model.prompt_learner.context_bank[n].grad for n in class_names(C)])
    assert (out["logits"] - gold["logits"]).abs().max().item() < 5e-5
    assert abs(out["loss"].item() - gold["loss"].item()) < 1e-5
    assert ((g - gold["ctx_grad"]).norm() / gold["ctx_grad"].norm()).item() < 1e-4
    assert abs(model.logit_scale.grad.item() - gold["logit_scale_grad"].item()) < 1e-5
    assert (model.last_attribution - gold["attribution"]).abs().max().item() < 1e-6
    assert set(model(images)) == {"logits"}                                        # model_wrapper.py:88-93


@pytest.mark.parametrize("mode", ["literal", "intended"])
@pytest.mark.parametrize("case", ["mini16_gate_b3c4p4", "minit512_resid_b2c3p5"])
def test_fullmodel_gate_and_residual_adjustors_match_reference_golden(case, mode):
    import tapclip_b200 as tb
    gold = load_golden(case, mode)
    B, C, P = gold["B"], gold["C"], gold["P"]
    ow, _ = build_oracle(gold["model_name"], C, P, mode)
    torch.manual_seed(4)
    model = tb.FullModel(class_names(C), FakeWrapper(ow, mode), prompt_len=P, adjustor_method=gold["method"])
    assert set(model.prompt_adjustor.state_dict()) == set(gold["adjustor_state"])
    for k, v in model.prompt_adjustor.state_dict().items():
        assert torch.equal(v, gold["adjustor_state"][k])                           # same constructor RNG order as the reference
    images, labels = synthetic_images(B, get_config(gold["model_name"]).image_size), synthetic_labels(B, C)
    model.train()
    out = model(images, labels)
    out["loss"].backward()
    g = torch.stack([model.prompt_learner.context_bank[n].grad for n in class_names(C)])
    assert (out["logits"] - gold["logits"]).abs().max().item() < 5e-5
    assert ((g - gold["ctx_grad"]).norm() / gold["ctx_grad"].norm()).item() < 1e-4
    for k, p in model.prompt_adjustor.named_parameters():
        ref = gold["adjustor_grad"][k]
        assert ((p.grad - ref).norm() / ref.norm().clamp_min(1e-12)).item() < 1e-4, k
    assert (model.last_attribution - gold["attribution"]).abs().max().item() < 1e-6


def test_flat_bank_parameter_identity_and_growth():
    gold, ow, om, model = _models("mini16_b4c5p4", "literal")
    pl = model.prompt_learner
    params = dict(pl.context_bank.items())
    before = {k: v.detach().clone() for k, v in params.items()}
    torch.manual_seed(7)
    for n in class_names(40)[5:]:
        pl.add_class_prompt(n)                                                     # forces several re-allocations
    assert pl.n_cls == 40 and pl.flat_ctx().shape == (40, gold["P"], 256) and pl.flat_ctx().is_contiguous()
    for i, (k, p) in enumerate(pl.context_bank.items()):
        assert p.data_ptr() == pl.flat_ctx()[i].data_ptr()                          # every Parameter is a row of the bank
        if k in params:
            assert p is params[k] and torch.equal(p.detach(), before[k])           # identity + value survive growth
    assert pl().shape == (40, gold["P"] + 77, 256)
    assert torch.equal(pl()[:, gold["P"]:], torch.cat([pl.token_bank[k] for k in pl.context_bank], 0))
    # an optimizer created BEFORE growth still updates the bank in place
    opt = torch.optim.SGD([params[class_names(5)[2]]], lr=1.0)
    params[class_names(5)[2]].grad = torch.ones_like(params[class_names(5)[2]])
    opt.step()
    assert torch.allclose(pl.flat_ctx()[2], before[class_names(5)[2]] - 1.0)
    # a Parameter re-pointed behind our back (e.g. p.data = ...) is re-packed on the next flat_ctx()
    params[class_names(5)[1]].data = torch.full((gold["P"], 256), 3.0)
    assert torch.equal(pl.flat_ctx()[1], torch.full((gold["P"], 256), 3.0))


def test_state_dict_surface():
    gold, ow, om, model = _models("mini16_b4c5p4", "intended")
    sd = model.state_dict()
    ref_sd_keys = set(om.state_dict().keys())
    assert {k for k in sd if k.startswith("prompt_learner.") or k == "logit_scale"} == \
           {k for k in ref_sd_keys if k.startswith("prompt_learner.") or k == "logit_scale"}
    assert sum("prompt_learner.context_bank" in n for n, _ in model.named_parameters()) == gold["C"]
    new = {k: torch.randn_like(v) for k, v in sd.items() if "context_bank" in k}
    model.load_state_dict(new, strict=False)                                       # test_cross_domain.py:61
    for i, n in enumerate(class_names(gold["C"])):
        assert torch.equal(model.prompt_learner.flat_ctx()[i], new[f"prompt_learner.context_bank.{n}"])


def test_prompt_adjustor_and_errors():
    import tapclip_b200 as tb
    with pytest.raises(ValueError, match="Unknown method"):
        tb.PromptAdjustor("bogus")
    x, a = torch.randn(2, 3, 8), torch.rand(2, 3)
    assert torch.equal(tb.PromptAdjustor("scale")(x, a), x * a.unsqueeze(-1))
    # 'gate' / 'residual' against the reference module itself (models/prompt_adjustor.py:13-25,38-44), same weights
    import importlib.util, os
    ref_path = "/root/reference/models/prompt_adjustor.py"
    if os.path.exists(ref_path):
        spec = importlib.util.spec_from_file_location("ref_prompt_adjustor", ref_path)
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        x = torch.randn(2, 3, 512)
        for method in ("gate", "residual"):
            torch.manual_seed(0)
            mine = tb.PromptAdjustor(method, dim=512)
            theirs = ref.PromptAdjustor(method)
            assert set(mine.state_dict()) == set(theirs.state_dict())
            theirs.load_state_dict(mine.state_dict())
            assert torch.equal(mine(x, a), theirs(x, a))


def test_class_sharding_bounds():
    from tapclip_b200.parallel import ClassSharding
    for world in (1, 2, 4, 8):
        for n in (1, 7, 65, 345):
            cover = []
            for r in range(world):
                lo, hi = ClassSharding(r, world).bounds(n)
                assert hi - lo <= ClassSharding(r, world).max_shard(n)
                cover += list(range(lo, hi))
            assert cover == list(range(n))


def test_bench_line_helpers():
    """bench.py pieces that need no GPU: both arms print the same `config` object, `roofline.traffic` is read from the committed
    ncu capture (not a literal), and the CPU legs time the same sub-grid."""
    import bench
    name, B, C, P, train, desc = bench.WORKLOADS["train_c2"]
    assert (name, B, C, P, train) == ("ViT-B-16-quickgelu", 128, 65, 16, True)          # BASELINE configs[1]
    a = bench.config_dict(name, B, C, P, train, desc, 1)
    assert set(a) == {"workload", "model", "batch_per_gpu", "global_batch", "n_cls", "prompt_len", "attribution", "optimizer", "parallelism"}
    assert bench.config_dict(name, B, C, P, train, desc, 8)["global_batch"] == 8 * B
    traffic, src = bench.ncu_gemm_traffic()
    assert src is not None and src.startswith("r02") and 50e6 < traffic < 400e6           # ~164 MB per image-tower GEMM launch
    assert bench.CPU_SAMPLE == (4, 8)
