"""Pins the CPU oracle (oracle/wan_dit.py) to the reference: golden vectors in tests/golden/ were produced by
the reference's OWN modules (oracle/make_golden.py).  bf16 runs must match bit for bit."""
import math
import os

import numpy as np
import pytest
import torch

import wan_dit
from conftest import GOLDEN

CONFIGS = {
    "dit_s72": dict(dim=256, heads=2, ffn_dim=512, num_layers=2, text_dim=64, freq_dim=256),
    "dit_s400": dict(dim=256, heads=2, ffn_dim=384, num_layers=1, text_dim=64, freq_dim=256),
}


def _load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def _inputs(g, dt):
    return (torch.from_numpy(g["latents"]).to(dt), torch.from_numpy(g["timestep"]), torch.from_numpy(g["text"]).to(dt))


@pytest.mark.parametrize("name", list(CONFIGS))
def test_dit_bf16_bit_exact_vs_reference(name):
    cfg, g = CONFIGS[name], _load(name)
    w = wan_dit.make_weights(**cfg, seed=1234, dtype=torch.bfloat16)
    lat, t, text = _inputs(g, torch.bfloat16)
    y = wan_dit.dit_forward(lat, t, text, w, heads=cfg["heads"], num_layers=cfg["num_layers"], freq_dim=cfg["freq_dim"])
    assert torch.equal(y.float(), torch.from_numpy(g["out_bf16"]))


@pytest.mark.parametrize("name", list(CONFIGS))
def test_dit_fp32_matches_reference_with_its_aliasing_quirk(name):
    """The reference's fp32 path squares q/k in place inside InplaceRMSNorm (x.float() aliases x); with that quirk
    reproduced the oracle is bit-exact in fp32 too, without it it is the exact-math version of the bf16 path."""
    cfg, g = CONFIGS[name], _load(name)
    w = wan_dit.make_weights(**cfg, seed=1234, dtype=torch.float32)
    lat, t, text = _inputs(g, torch.float32)
    kw = dict(heads=cfg["heads"], num_layers=cfg["num_layers"], freq_dim=cfg["freq_dim"])
    wan_dit.REF_FP32_ALIAS_QUIRK = True
    try:
        y = wan_dit.dit_forward(lat, t, text, w, **kw)
    finally:
        wan_dit.REF_FP32_ALIAS_QUIRK = False
    assert torch.equal(y, torch.from_numpy(g["out_fp32"]))
    # exact-math oracle vs the reference's bf16 output: bf16-level agreement
    y32 = wan_dit.dit_forward(lat, t, text, w, **kw)
    ref16 = torch.from_numpy(g["out_bf16"])
    rel = ((y32 - ref16).norm() / y32.norm()).item()
    assert rel < 2e-2, rel


@pytest.mark.parametrize("name", list(CONFIGS))
def test_intermediates(name):
    cfg, g = CONFIGS[name], _load(name)
    for dt, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        w = wan_dit.make_weights(**cfg, seed=1234, dtype=dt)
        lat, t, text = _inputs(g, dt)
        temb, temb6, ctx = wan_dit.condition_embed(t, text, w, cfg["freq_dim"])
        assert torch.equal(temb.float(), torch.from_numpy(g["temb_" + tag]))
        assert torch.equal(temb6.flatten(1).float(), torch.from_numpy(g["tproj_" + tag]))
        assert torch.equal(ctx.float(), torch.from_numpy(g["ctx_" + tag]))
        h = wan_dit.patchify(lat, w)
        assert torch.equal(h.float(), torch.from_numpy(g["patch_" + tag]))
        grid = (lat.shape[2], lat.shape[3] // 2, lat.shape[4] // 2)
        fr = wan_dit.rope_table(128, grid)
        assert np.array_equal(fr.real.numpy(), g["rope_real"]) and np.array_equal(fr.imag.numpy(), g["rope_imag"])
        if tag == "bf16":
            b0 = wan_dit.block_forward(h, ctx, temb6, fr, w, "blocks.0", cfg["heads"])
            assert torch.equal(b0.float(), torch.from_numpy(g["block0_bf16"]))


def test_product_rope_table_matches_reference_cast():
    """apex-studio_b200/wan/rope.py (product, host side) == reference table cast to bf16 (ops.py:157-158)."""
    from apex_studio_b200.wan.rope import wan_rope_table_bf16

    g = _load("dit_s400")
    tab = wan_rope_table_bf16(128, (5, 8, 10), "cpu").float().view(400, 64, 2)
    assert torch.equal(tab[..., 0], torch.from_numpy(g["rope_real"]).to(torch.bfloat16).float())
    assert torch.equal(tab[..., 1], torch.from_numpy(g["rope_imag"]).to(torch.bfloat16).float())


def test_attention_recipe_of_the_reference():
    """scripts/smoke_tests/test_attention_backends.py:232-388 on CPU: 1x32x1024x128, seed 42, gold = `sdpa`,
    tolerance 1e-4 (fp32).  The oracle's exact-math attention must meet it."""
    g = _load("attention")
    torch.manual_seed(42)
    q, k, v = torch.randn(1, 32, 1024, 128), torch.randn(1, 32, 1024, 128), torch.randn(1, 32, 1024, 128)
    heads = g["heads"].tolist()
    out = wan_dit.sdpa_fp32_math(q[:, heads], k[:, heads], v[:, heads])
    gold = torch.from_numpy(g["gold_1x32x1024x128_seed42"])
    max_abs = (out - gold).abs().max().item()
    assert max_abs <= 1e-4 or max_abs / gold.abs().max().item() <= 1e-4
    out2 = wan_dit.sdpa_fp32_math(torch.from_numpy(g["q2"]), torch.from_numpy(g["k2"]), torch.from_numpy(g["v2"]))
    assert (out2 - torch.from_numpy(g["gold2"])).abs().max().item() <= 1e-4


def test_cfg_combine_and_rounding_points():
    c = torch.randn(1000).bfloat16()
    u = torch.randn(1000).bfloat16()
    r = wan_dit.cfg_combine(c, u, 4.0)
    expect = (u.float() + (4.0 * (c.float() - u.float()).bfloat16().float()).bfloat16().float()).bfloat16()
    assert torch.equal(r, expect)
