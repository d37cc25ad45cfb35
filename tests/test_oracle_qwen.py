"""Pins the QwenImage CPU oracle (oracle/qwen_dit.py) to the reference: tests/golden/qwen_*.npz were produced by the
reference's OWN QwenImageTransformer2DModel (oracle/make_golden.py golden_qwen); fp32 and bf16 must match bit for bit."""
import os

import numpy as np
import pytest
import torch

import qwen_dit
from conftest import GOLDEN

CONFIGS = {
    "qwen_t2i": dict(dim=256, heads=2, num_layers=2, in_channels=16, out_channels=4, joint_dim=48),
    "qwen_edit": dict(dim=256, heads=2, num_layers=1, in_channels=16, out_channels=4, joint_dim=48),
}


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def inputs(g, dt):
    t = lambda k: torch.from_numpy(g[k])
    shapes = [tuple(int(v) for v in row) for row in g["img_shapes"]]
    return t("hidden").to(dt), t("enc").to(dt), t("timestep").to(dt), shapes, int(g["enc"].shape[1])


def kw(cfg):
    return dict(heads=cfg["heads"], num_layers=cfg["num_layers"])


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("tag,dt", [("fp32", torch.float32), ("bf16", torch.bfloat16)])
def test_forward_bit_exact_vs_reference(name, tag, dt):
    cfg, g = CONFIGS[name], load(name)
    w = qwen_dit.make_weights(**cfg, seed=1234, dtype=dt)
    y = qwen_dit.qwen_forward(*inputs(g, dt), w, **kw(cfg))
    assert torch.equal(y.float(), torch.from_numpy(g["out_" + tag]))


@pytest.mark.parametrize("name", list(CONFIGS))
def test_intermediates_bit_exact(name):
    cfg, g = CONFIGS[name], load(name)
    for tag, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        w = qwen_dit.make_weights(**cfg, seed=1234, dtype=dt)
        x, enc, t, shapes, n_txt = inputs(g, dt)
        temb = qwen_dit.time_embed(t, w, dt)
        assert torch.equal(temb.float(), torch.from_numpy(g["temb_" + tag]))
        vf, tf = qwen_dit.rope_tables(shapes, n_txt)
        assert torch.equal(vf.real, torch.from_numpy(g["img_freqs_re"])) and torch.equal(vf.imag, torch.from_numpy(g["img_freqs_im"]))
        assert torch.equal(tf.real, torch.from_numpy(g["txt_freqs_re"])) and torch.equal(tf.imag, torch.from_numpy(g["txt_freqs_im"]))
        ctx = qwen_dit.linear(qwen_dit.rms_norm(enc, w["txt_norm.weight"]), w, "txt_in")
        assert torch.equal(ctx.float(), torch.from_numpy(g["ctx_in_" + tag]))
        c1, h1 = qwen_dit.dual_block(0, w, cfg["heads"], qwen_dit.linear(x, w, "img_in"), ctx, temb, vf, tf)
        assert torch.equal(h1.float(), torch.from_numpy(g["block0_x_" + tag]))
        assert torch.equal(c1.float(), torch.from_numpy(g["block0_ctx_" + tag]))


# ------------------------------------------------------------------------------------------------ host mirror, no GPU
def _product(cfg):
    from apex_studio_b200.qwenimage import QwenImageConfig, QwenImageTransformer2DModel

    return QwenImageTransformer2DModel(QwenImageConfig(
        in_channels=cfg["in_channels"], out_channels=cfg["out_channels"], num_layers=cfg["num_layers"],
        num_attention_heads=cfg["heads"], joint_attention_dim=cfg["joint_dim"]))


def test_product_state_dict_fusion_and_rope_tables():
    from apex_studio_b200.qwenimage import qwen_rope_tables

    cfg = CONFIGS["qwen_edit"]
    w = qwen_dit.make_weights(**cfg, seed=1234)
    m = _product(cfg)
    assert set(m.state_dict_keys()) == set(w)
    m.load_state_dict(w, device="cpu")
    d = 256
    assert m.w["modulation.weight"].shape == (12 * d + 2 * d, d)
    assert m.lora_target("transformer_blocks.0.txt_mod.1")[:3] == ("modulation.weight", 6 * d, 6 * d)
    assert m.lora_target("transformer_blocks.0.attn.add_k_proj")[:3] == ("transformer_blocks.0.attn.add_qkv.weight", d, d)
    assert m.w["proj_out.weight"].shape == (16, d) and m.lora_target("proj_out")[:3] == ("proj_out.weight", 0, 16)
    g = load("qwen_edit")
    _, _, _, shapes, n_txt = inputs(g, torch.float32)
    img, txt = qwen_rope_tables(shapes, n_txt, (16, 56, 56), "cpu")
    assert img.shape == (156, 64, 2) and txt.shape == (13, 64, 2)
    assert torch.equal(img[..., 0], torch.from_numpy(g["img_freqs_re"])) and torch.equal(img[..., 1], torch.from_numpy(g["img_freqs_im"]))
    assert torch.equal(txt[..., 0], torch.from_numpy(g["txt_freqs_re"])) and torch.equal(txt[..., 1], torch.from_numpy(g["txt_freqs_im"]))
    from apex_studio_b200.qwenimage import QwenImageConfig, QwenImageTransformer2DModel
    with pytest.raises(ValueError):
        QwenImageTransformer2DModel(QwenImageConfig(zero_cond_t=True))
    x, enc, t, shapes, n_txt = inputs(g, torch.bfloat16)
    with pytest.raises(ValueError):            # no CPU fallback
        m(hidden_states=x, encoder_hidden_states=enc, timestep=t, img_shapes=[shapes], txt_seq_lens=[n_txt])
