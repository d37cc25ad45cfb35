"""Pins the Flux CPU oracle (oracle/flux_dit.py) to the reference: tests/golden/flux_*.npz were produced by the
reference's OWN FluxTransformer2DModel (oracle/make_golden.py golden_flux).  fp32 and bf16 runs must match bit for bit.
Also host-logic tests of the Flux mirror that need no GPU (state-dict fusion, LoRA row addressing, scheduler, packing)."""
import os

import numpy as np
import pytest
import torch

import flux_dit
from conftest import GOLDEN

CONFIGS = {
    "flux_s56": dict(dim=256, heads=2, num_layers=2, num_single_layers=2, in_channels=16, joint_dim=32, pooled_dim=24,
                     guidance_embeds=True),
    "flux_s200": dict(dim=256, heads=2, num_layers=1, num_single_layers=1, in_channels=16, joint_dim=32, pooled_dim=24,
                      guidance_embeds=False),
}


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def inputs(g, dt):
    t = lambda k: torch.from_numpy(g[k])
    guidance = t("guidance") if "guidance" in g.files else None
    return (t("hidden").to(dt), t("enc").to(dt), t("pooled").to(dt), t("timestep").to(dt), t("img_ids"), t("txt_ids"), guidance)


def kw(cfg):
    return dict(heads=cfg["heads"], num_layers=cfg["num_layers"], num_single_layers=cfg["num_single_layers"])


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("tag,dt", [("fp32", torch.float32), ("bf16", torch.bfloat16)])
def test_forward_bit_exact_vs_reference(name, tag, dt):
    cfg, g = CONFIGS[name], load(name)
    w = flux_dit.make_weights(**cfg, seed=1234, dtype=dt)
    y = flux_dit.flux_forward(*inputs(g, dt), w, **kw(cfg))
    assert torch.equal(y.float(), torch.from_numpy(g["out_" + tag]))


@pytest.mark.parametrize("name", list(CONFIGS))
def test_intermediates_bit_exact(name):
    cfg, g = CONFIGS[name], load(name)
    for tag, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        w = flux_dit.make_weights(**cfg, seed=1234, dtype=dt)
        x, enc, pooled, t, img_ids, txt_ids, guidance = inputs(g, dt)
        gd = guidance.to(dt) * 1000 if guidance is not None else None
        temb = flux_dit.time_text_embed(t * 1000, gd, pooled, w)
        assert torch.equal(temb.float(), torch.from_numpy(g["temb_" + tag]))
        cos, sin = flux_dit.rope_table(torch.cat((txt_ids, img_ids), dim=0))
        assert torch.equal(cos, torch.from_numpy(g["rope_cos"])) and torch.equal(sin, torch.from_numpy(g["rope_sin"]))
        ctx, h = flux_dit.linear(enc, w, "context_embedder"), flux_dit.linear(x, w, "x_embedder")
        ctx, h = flux_dit.dual_block(0, w, cfg["heads"], h, ctx, temb, (cos, sin))
        assert torch.equal(h.float(), torch.from_numpy(g["dual0_x_" + tag]))
        assert torch.equal(ctx.float(), torch.from_numpy(g["dual0_ctx_" + tag]))
        ctx, h = flux_dit.single_block(0, w, cfg["heads"], h, ctx, temb, (cos, sin))
        assert torch.equal(h.float(), torch.from_numpy(g["single0_x_" + tag]))
        assert torch.equal(ctx.float(), torch.from_numpy(g["single0_ctx_" + tag]))


def test_text_rows_of_the_rope_table_are_the_identity():
    """txt_ids are zeros in every Flux engine (engine/flux/shared.py), so cos = 1 and sin = 0 exactly on the text rows:
    the product skips the rotation there."""
    g = load("flux_s56")
    n_txt = g["txt_ids"].shape[0]
    assert (g["rope_cos"][:n_txt] == 1.0).all() and (g["rope_sin"][:n_txt] == 0.0).all()


def test_pack_unpack_and_ids():
    lat = torch.randn(2, 16, 8, 12)
    packed = flux_dit.pack_latents(lat)
    assert packed.shape == (2, 24, 64)
    assert torch.equal(flux_dit.unpack_latents(packed, 64, 96), lat)
    ids = flux_dit.latent_image_ids(4, 6)
    assert ids.shape == (24, 3) and ids[7].tolist() == [0.0, 1.0, 1.0]


# ------------------------------------------------------------------------------------------------ host mirror, no GPU
def test_product_rope_table_matches_reference_table():
    from apex_studio_b200.flux import flux_rope_table

    g = load("flux_s200")
    ids = torch.cat((torch.from_numpy(g["txt_ids"]), torch.from_numpy(g["img_ids"])), dim=0)
    tab = flux_rope_table(ids, (16, 56, 56), "cpu")
    assert tab.shape == (200, 64, 2) and tab.dtype == torch.float32
    assert torch.equal(tab[..., 0], torch.from_numpy(g["rope_cos"])[:, ::2])
    assert torch.equal(tab[..., 1], torch.from_numpy(g["rope_sin"])[:, 1::2])


def test_state_dict_fusion_and_lora_rows_on_cpu():
    from apex_studio_b200.flux import FluxConfig, FluxTransformer2DModel

    cfg = CONFIGS["flux_s56"]
    w = flux_dit.make_weights(**cfg, seed=1234)
    m = FluxTransformer2DModel(FluxConfig(in_channels=16, num_layers=2, num_single_layers=2, num_attention_heads=2,
                                          joint_attention_dim=32, pooled_projection_dim=24, guidance_embeds=True))
    assert set(m.state_dict_keys()) == set(w)
    m.load_state_dict(w, device="cpu")
    d = 256
    assert m.w["transformer_blocks.1.attn.to_qkv.weight"].shape == (3 * d, d)
    assert torch.equal(m.w["transformer_blocks.1.attn.add_qkv.weight"][d:2 * d], w["transformer_blocks.1.attn.add_k_proj.weight"].bfloat16())
    total = 2 * 12 * d + 2 * 3 * d + 2 * d
    assert m.w["modulation.weight"].shape == (total, d)
    key, r0, rows, bkey = m.lora_target("single_transformer_blocks.1.norm.linear")
    assert (key, rows, bkey) == ("modulation.weight", 3 * d, "modulation.bias") and r0 == 2 * 12 * d + 3 * d
    assert torch.equal(m.w[key][r0:r0 + rows], w["single_transformer_blocks.1.norm.linear.weight"].bfloat16())
    assert m.lora_target("transformer_blocks.0.attn.to_v")[:3] == ("transformer_blocks.0.attn.to_qkv.weight", 2 * d, d)
    assert m.lora_target("transformer_blocks.0.attn.add_q_proj")[:3] == ("transformer_blocks.0.attn.add_qkv.weight", 0, d)
    assert m.lora_target("single_transformer_blocks.0.proj_out")[:3] == ("single_transformer_blocks.0.proj_out.weight", 0, d)
    with pytest.raises(ValueError):
        m.lora_target("transformer_blocks.0.attn.nope")
    with pytest.raises(KeyError):
        FluxTransformer2DModel(FluxConfig(in_channels=16, num_layers=1, num_single_layers=2, num_attention_heads=2,
                                          joint_attention_dim=32, pooled_projection_dim=24)).load_state_dict(w, device="cpu")
    with pytest.raises(ValueError):                      # no CPU fallback: the forward needs the CUDA kernels
        x, enc, pooled, t, img_ids, txt_ids, guidance = inputs(load("flux_s56"), torch.bfloat16)
        m(x, enc, pooled, t, img_ids, txt_ids, guidance)


def test_flow_match_euler_product_equals_oracle_and_engine_call_pattern():
    from apex_studio_b200.scheduler import FlowMatchEulerDiscreteScheduler, calculate_shift

    mu = calculate_shift(4096)
    assert mu == flux_dit.calculate_shift(4096) == pytest.approx(1.15)
    assert calculate_shift(256) == pytest.approx(0.5)
    o, s = flux_dit.FlowMatchEuler(), FlowMatchEulerDiscreteScheduler()
    n = 28
    to = o.set_timesteps(n, mu)
    ts = s.set_timesteps(n, sigmas=np.linspace(1.0, 1 / n, n), mu=mu)      # engine/flux/t2i.py:110-135
    assert torch.equal(to, ts) and ts[0] == 1000.0 and len(ts) == n and s.sigmas[-1] == 0
    assert bool((ts[1:] < ts[:-1]).all())
    x = torch.randn(1, 16, 64, generator=torch.Generator().manual_seed(0)).bfloat16()
    xs = x.clone()
    for i, t in enumerate(ts):
        mo = torch.randn(1, 16, 64, generator=torch.Generator().manual_seed(i)).bfloat16()
        x, xs = o.step(mo, t, x), s.step(mo, t, xs)[0]
    assert torch.equal(x, xs) and s.step_index == n
    with pytest.raises(ValueError):
        s.set_timesteps(4)                                # dynamic shifting needs mu
