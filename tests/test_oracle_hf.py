"""Independent cross-check of the restated open_clip model: HuggingFace `transformers` CLIP, after remapping the
weights (packed in_proj <-> separate q/k/v), must reproduce the stand-in's towers, per-layer attention
probabilities and the standard encode_text path.  (open_clip itself is absent: SURVEY 8c — "parity unpinned" for
the model, pinned here against a second implementation.)"""
import pytest
import torch

from oracle.clip_standin import build_clip, vision_cls_attention

transformers = pytest.importorskip("transformers")


def _remap(src, hf_model, cfg):
    sd = src.state_dict()
    out = {}

    def blocks(prefix_src, prefix_hf, n, d):
        for i in range(n):
            s, h = f"{prefix_src}.{i}.", f"{prefix_hf}.{i}."
            w, b = sd[s + "attn.in_proj_weight"], sd[s + "attn.in_proj_bias"]
            for j, nm in enumerate(("q_proj", "k_proj", "v_proj")):
                out[h + f"self_attn.{nm}.weight"] = w[j * d:(j + 1) * d]
                out[h + f"self_attn.{nm}.bias"] = b[j * d:(j + 1) * d]
            out[h + "self_attn.out_proj.weight"] = sd[s + "attn.out_proj.weight"]
            out[h + "self_attn.out_proj.bias"] = sd[s + "attn.out_proj.bias"]
            for a, bname in (("ln_1", "layer_norm1"), ("ln_2", "layer_norm2")):
                out[h + bname + ".weight"], out[h + bname + ".bias"] = sd[s + a + ".weight"], sd[s + a + ".bias"]
            for a, bname in (("mlp.c_fc", "mlp.fc1"), ("mlp.c_proj", "mlp.fc2")):
                out[h + bname + ".weight"], out[h + bname + ".bias"] = sd[s + a + ".weight"], sd[s + a + ".bias"]

    blocks("visual.transformer.resblocks", "vision_model.encoder.layers", cfg.vision_layers, cfg.vision_width)
    blocks("transformer.resblocks", "text_model.encoder.layers", cfg.text_layers, cfg.text_width)
    out["vision_model.embeddings.class_embedding"] = sd["visual.class_embedding"]
    out["vision_model.embeddings.patch_embedding.weight"] = sd["visual.conv1.weight"]
    out["vision_model.embeddings.position_embedding.weight"] = sd["visual.positional_embedding"]
    out["vision_model.pre_layrnorm.weight"], out["vision_model.pre_layrnorm.bias"] = sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"]
    out["vision_model.post_layernorm.weight"], out["vision_model.post_layernorm.bias"] = sd["visual.ln_post.weight"], sd["visual.ln_post.bias"]
    out["visual_projection.weight"] = sd["visual.proj"].t()
    out["text_model.embeddings.token_embedding.weight"] = sd["token_embedding.weight"]
    out["text_model.embeddings.position_embedding.weight"] = sd["positional_embedding"]
    out["text_model.final_layer_norm.weight"], out["text_model.final_layer_norm.bias"] = sd["ln_final.weight"], sd["ln_final.bias"]
    out["text_projection.weight"] = sd["text_projection"].t()
    out["logit_scale"] = sd["logit_scale"]
    missing, unexpected = hf_model.load_state_dict(out, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in m for m in missing), missing


@pytest.mark.parametrize("name,act", [("mini-16-quickgelu", "quick_gelu"), ("mini-14", "gelu")])
def test_standin_towers_match_hf_clip(name, act):
    from transformers import CLIPConfig, CLIPModel
    src = build_clip(name, seed=0)
    c = src.cfg
    hf_cfg = CLIPConfig(
        text_config=dict(hidden_size=c.text_width, intermediate_size=4 * c.text_width, num_hidden_layers=c.text_layers,
                         num_attention_heads=c.text_heads, max_position_embeddings=c.context_length, vocab_size=c.vocab_size,
                         hidden_act=act, projection_dim=c.embed_dim, eos_token_id=49407, bos_token_id=49406, pad_token_id=0),
        vision_config=dict(hidden_size=c.vision_width, intermediate_size=4 * c.vision_width, num_hidden_layers=c.vision_layers,
                           num_attention_heads=c.vision_heads, image_size=c.image_size, patch_size=c.patch_size, hidden_act=act,
                           projection_dim=c.embed_dim),
        projection_dim=c.embed_dim)
    hf_cfg._attn_implementation = "eager"
    hf = CLIPModel(hf_cfg).eval()
    _remap(src, hf, c)
    g = torch.Generator().manual_seed(0)
    images = torch.randn(3, 3, c.image_size, c.image_size, generator=g)
    with torch.no_grad():
        feats, cls_rows = vision_cls_attention(src, images)
        vout = hf.vision_model(pixel_values=images, output_attentions=True)
        hf_feats = hf.visual_projection(vout.pooler_output)
        assert (feats - hf_feats).abs().max().item() < 2e-4
        hf_rows = torch.stack([a[:, :, 0, :] for a in vout.attentions], dim=1)
        assert (cls_rows - hf_rows).abs().max().item() < 1e-5
        # standard text path: positional embedding + causal mask + ln_final + EOT pooling
        from oracle.clip_standin import SyntheticTokenizer
        tok = SyntheticTokenizer(c.context_length)(["a photo of a cat", "a photo of a tall giraffe"])
        ref = src.encode_text(tok)
        tout = hf.text_model(input_ids=tok)
        hf_text = hf.text_projection(tout.pooler_output)
        assert (ref - hf_text).abs().max().item() < 2e-4
