"""Pins the Flux2 CPU oracle (oracle/flux2_dit.py) to the reference: tests/golden/flux2_*.npz were produced by the
reference's OWN Flux2Transformer2DModel (oracle/make_golden.py golden_flux2); fp32 and bf16 must match bit for bit."""
import os

import numpy as np
import pytest
import torch

import flux2_dit
from conftest import GOLDEN

CONFIGS = {
    "flux2_s56": dict(dim=256, heads=2, num_layers=2, num_single_layers=2, in_channels=16, joint_dim=48, guidance_embeds=False),
    "flux2_s200": dict(dim=256, heads=2, num_layers=1, num_single_layers=1, in_channels=16, joint_dim=48, guidance_embeds=True),
}


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def inputs(g, dt):
    t = lambda k: torch.from_numpy(g[k])
    guidance = t("guidance").to(dt) if "guidance" in g.files else None
    return t("hidden").to(dt), t("enc").to(dt), t("timestep").to(dt), t("img_ids"), t("txt_ids"), guidance


def kw(cfg):
    return dict(heads=cfg["heads"], num_layers=cfg["num_layers"], num_single_layers=cfg["num_single_layers"])


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("tag,dt", [("fp32", torch.float32), ("bf16", torch.bfloat16)])
def test_forward_bit_exact_vs_reference(name, tag, dt):
    cfg, g = CONFIGS[name], load(name)
    w = flux2_dit.make_weights(**cfg, seed=1234, dtype=dt)
    y = flux2_dit.flux2_forward(*inputs(g, dt), w, **kw(cfg))
    assert torch.equal(y.float(), torch.from_numpy(g["out_" + tag]))


def _product(cfg):
    from apex_studio_b200.flux2 import Flux2Config, Flux2Transformer2DModel

    return Flux2Transformer2DModel(Flux2Config(in_channels=cfg["in_channels"], num_layers=cfg["num_layers"],
                                               num_single_layers=cfg["num_single_layers"], num_attention_heads=cfg["heads"],
                                               joint_attention_dim=cfg["joint_dim"], guidance_embeds=cfg["guidance_embeds"]))


def test_product_state_dict_and_lora_rows_on_cpu():
    cfg = CONFIGS["flux2_s200"]
    w = flux2_dit.make_weights(**cfg, seed=1234)
    m = _product(cfg)
    assert set(m.state_dict_keys()) == set(w)
    m.load_state_dict(w, device="cpu")
    d = 256
    assert m.w["transformer_blocks.0.attn.add_qkv.weight"].shape == (3 * d, d) and m.mlp == 768
    assert m.lora_target("transformer_blocks.0.attn.to_k")[:3] == ("transformer_blocks.0.attn.to_qkv.weight", d, d)
    assert m.lora_target("single_transformer_blocks.0.attn.to_qkv_mlp_proj")[:3] == (
        "single_transformer_blocks.0.attn.to_qkv_mlp_proj.weight", 0, 3 * d + 2 * 768)
    x, enc, t, img_ids, txt_ids, guidance = inputs(load("flux2_s200"), torch.bfloat16)
    with pytest.raises(ValueError):            # no CPU fallback
        m(x, enc, t, img_ids, txt_ids, guidance)
