"""Pins the HunyuanVideo-1.5 CPU oracle (oracle/hy15_dit.py) to the reference: tests/golden/hy15_*.npz were produced by the
reference's OWN HunyuanVideo15Transformer3DModel (oracle/make_golden.py golden_hy15).  bf16 must match bit for bit; fp32
matches bit for bit once the reference's fp32 aliasing quirk of InplaceRMSNorm is switched on in the oracle."""
import os

import numpy as np
import pytest
import torch

import hy15_dit
import wan_dit
from conftest import GOLDEN

CONFIGS = {
    "hy15_t2v": dict(dim=256, heads=2, num_layers=2, num_refiner_layers=2, in_channels=9, out_channels=4, text_dim=48,
                     text2_dim=40, image_dim=24, byt5_hidden=64),
    "hy15_i2v": dict(dim=256, heads=2, num_layers=1, num_refiner_layers=1, in_channels=9, out_channels=4, text_dim=48,
                     text2_dim=40, image_dim=24, byt5_hidden=64),
}


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def inputs(g, dt):
    t = lambda k: torch.from_numpy(g[k])
    return (t("hidden").to(dt), t("timestep").to(dt), t("text").to(dt), t("mask"), t("text2").to(dt), t("mask2"), t("image").to(dt))


def kw(cfg):
    return dict(heads=cfg["heads"], num_layers=cfg["num_layers"], num_refiner_layers=cfg["num_refiner_layers"])


@pytest.mark.parametrize("name", list(CONFIGS))
def test_forward_bf16_bit_exact_vs_reference(name):
    cfg, g = CONFIGS[name], load(name)
    w = hy15_dit.make_weights(**cfg, seed=1234, dtype=torch.bfloat16)
    y = hy15_dit.hy15_forward(*inputs(g, torch.bfloat16), w, **kw(cfg))
    assert torch.equal(y.float(), torch.from_numpy(g["out_bf16"]))


@pytest.mark.parametrize("name", list(CONFIGS))
def test_forward_fp32_matches_reference_with_its_aliasing_quirk(name):
    cfg, g = CONFIGS[name], load(name)
    w = hy15_dit.make_weights(**cfg, seed=1234, dtype=torch.float32)
    wan_dit.REF_FP32_ALIAS_QUIRK = True
    try:
        y = hy15_dit.hy15_forward(*inputs(g, torch.float32), w, **kw(cfg))
    finally:
        wan_dit.REF_FP32_ALIAS_QUIRK = False
    assert torch.equal(y, torch.from_numpy(g["out_fp32"]))
    exact = hy15_dit.hy15_forward(*inputs(g, torch.float32), w, **kw(cfg))       # exact-math version of the bf16 path
    ref16 = torch.from_numpy(g["out_bf16"])
    assert ((exact - ref16).norm() / exact.norm()).item() < 3e-2


@pytest.mark.parametrize("name", list(CONFIGS))
def test_intermediates(name):
    cfg, g = CONFIGS[name], load(name)
    for tag, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        w = hy15_dit.make_weights(**cfg, seed=1234, dtype=dt)
        x, t, text, mask, text2, mask2, img = inputs(g, dt)
        assert torch.equal(hy15_dit.time_embed(t, w).float(), torch.from_numpy(g["temb_" + tag]))
        ref = hy15_dit.token_refiner(text, t, mask, w, cfg["heads"], cfg["num_refiner_layers"])
        valid = mask[0].bool()
        assert torch.equal(ref[0][valid].float(), torch.from_numpy(g["refined_" + tag])[0][valid])
        f, hh, ww = x.shape[2:]
        cos, sin = hy15_dit.rope_table((f, hh, ww))
        assert torch.equal(cos, torch.from_numpy(g["rope_cos"])) and torch.equal(sin, torch.from_numpy(g["rope_sin"]))


def test_refiner_on_compacted_valid_tokens_equals_masked_refiner():
    """The product runs the token refiner on the VALID tokens only (no mask): for the valid rows this equals the reference's
    key-padding-mask formulation; the padded rows are replaced by zeros in the reorder anyway (model.py:1068-1075)."""
    name = "hy15_t2v"
    cfg, g = CONFIGS[name], load(name)
    w = hy15_dit.make_weights(**cfg, seed=1234, dtype=torch.float32)
    x, t, text, mask, text2, mask2, img = inputs(g, torch.float32)
    full = hy15_dit.token_refiner(text, t, mask, w, cfg["heads"], cfg["num_refiner_layers"])
    valid = mask[0].bool()
    # pooled projection must still see the mask -> pass the compacted tokens with an all-ones mask of the valid length
    comp = hy15_dit.token_refiner(text[:, valid], t, torch.ones(1, int(valid.sum())), w, cfg["heads"], cfg["num_refiner_layers"])
    assert torch.allclose(full[0][valid], comp[0], atol=2e-6, rtol=1e-5)


# ------------------------------------------------------------------------------------------------ host mirror, no GPU
def _product(cfg):
    from apex_studio_b200.hunyuanvideo15 import HunyuanVideo15Config, HunyuanVideo15Transformer3DModel

    return HunyuanVideo15Transformer3DModel(HunyuanVideo15Config(
        in_channels=cfg["in_channels"], out_channels=cfg["out_channels"], num_attention_heads=cfg["heads"],
        num_layers=cfg["num_layers"], num_refiner_layers=cfg["num_refiner_layers"], text_embed_dim=cfg["text_dim"],
        text_embed_2_dim=cfg["text2_dim"], image_embed_dim=cfg["image_dim"]))


def test_product_state_dict_keys_fusion_and_rope_table():
    from apex_studio_b200.hunyuanvideo15 import hy15_rope_table

    cfg = CONFIGS["hy15_t2v"]
    w = hy15_dit.make_weights(**cfg, seed=1234)
    m = _product(cfg)
    assert set(m.state_dict_keys()) == set(w)
    m.load_state_dict(w, device="cpu")
    d = 256
    assert m.w["x_embedder.proj.weight"].shape == (d, 16) and m._k_pad == 7          # K = 9 padded to 16 for TMA
    assert torch.equal(m.w["x_embedder.proj.weight"][:, 9:], torch.zeros(d, 7, dtype=torch.bfloat16))
    assert m.w["transformer_blocks.1.attn.add_qkv.weight"].shape == (3 * d, d)
    assert m.w["context_embedder.token_refiner.refiner_blocks.0.attn.to_qkv.bias"].shape == (3 * d,)
    assert m.w["modulation.weight"].shape == (2 * 12 * d + 2 * d, d)
    key, r0, rows, _ = m.lora_target("transformer_blocks.1.norm1_context.linear")
    assert (key, r0, rows) == ("modulation.weight", 12 * d + 6 * d, 6 * d)
    assert m.lora_target("transformer_blocks.0.attn.add_v_proj")[:3] == ("transformer_blocks.0.attn.add_qkv.weight", 2 * d, d)
    with pytest.raises(ValueError):
        m.lora_target("x_embedder.proj")
    # rope table = the reference's table after its bf16 cast, one entry per channel pair
    g = load("hy15_i2v")
    tab = hy15_rope_table((5, 8, 10), (16, 56, 56), 256.0, "cpu")
    assert tab.shape == (400, 64, 2)
    assert torch.equal(tab[..., 0], torch.from_numpy(g["rope_cos"]).bfloat16().float()[:, ::2])
    assert torch.equal(tab[..., 1], torch.from_numpy(g["rope_sin"]).bfloat16().float()[:, ::2])
    # patchify order == Conv3d(kernel = stride = patch) im2col order
    x = torch.randn(9, 3, 4, 6)
    tok = m.patchify(x.bfloat16())[:, :9].float()
    ref = torch.nn.functional.conv3d(x.bfloat16().float()[None], torch.eye(9).view(9, 9, 1, 1, 1)).flatten(2).transpose(1, 2)[0]
    assert torch.equal(tok, ref)
    with pytest.raises(ValueError):            # no CPU fallback
        m(*[t for t in inputs(load("hy15_t2v"), torch.bfloat16)][:4], encoder_hidden_states_2=torch.zeros(1, 6, 40),
          encoder_attention_mask_2=torch.ones(1, 6), image_embeds=torch.zeros(1, 5, 24))
