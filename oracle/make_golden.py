"""Generate tests/golden/*.npz by running the REFERENCE's own modules (imported unmodified from
/root/reference through oracle/ref_import) on seeded synthetic inputs.  Run in the build container only:

    python oracle/make_golden.py            # writes tests/golden/*.npz

The fixtures pin (a) the CPU oracle restatement (tests/test_oracle_golden.py, CPU) and (b) the CUDA path
(tests/test_gpu_parity.py, GPU) to the reference.  bf16 tensors are stored widened to float32 (exact).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)

from ref_import import bootstrap  # noqa: E402
import wan_dit  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

DIT_CONFIGS = {
    # name: (model kwargs, latent shape, text len)
    "dit_s72": (dict(dim=256, heads=2, ffn_dim=512, num_layers=2, text_dim=64, freq_dim=256), (1, 16, 3, 8, 12), 16),
    "dit_s400": (dict(dim=256, heads=2, ffn_dim=384, num_layers=1, text_dim=64, freq_dim=256), (1, 16, 5, 16, 20), 40),
}


def f32(t: torch.Tensor) -> np.ndarray:
    return t.detach().float().cpu().numpy()


def build_reference_dit(cfg, weights):
    m = bootstrap.ref("src.transformer.wan.base.model")
    a = bootstrap.ref("src.attention.functions")
    a.attention_register.set_default("sdpa")
    model = m.WanTransformer3DModel(
        num_attention_heads=cfg["heads"], attention_head_dim=cfg["dim"] // cfg["heads"], in_channels=16,
        out_channels=16, text_dim=cfg["text_dim"], freq_dim=cfg["freq_dim"], ffn_dim=cfg["ffn_dim"],
        num_layers=cfg["num_layers"])
    missing, unexpected = model.load_state_dict(weights, strict=False)
    assert not unexpected, unexpected
    # only buffers-free modules may be "missing" (none expected)
    assert not missing, missing
    return model.eval()


def golden_dit():
    for name, (cfg, lshape, tlen) in DIT_CONFIGS.items():
        out = {}
        w32 = wan_dit.make_weights(**cfg, seed=1234, dtype=torch.float32)
        g = torch.Generator().manual_seed(42)
        latents = torch.randn(lshape, generator=g)
        text = torch.randn(1, tlen, cfg["text_dim"], generator=torch.Generator().manual_seed(43))
        t = torch.tensor([875], dtype=torch.int64)
        out["latents"], out["text"], out["timestep"] = f32(latents), f32(text), t.numpy()
        for dt, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
            model = build_reference_dit(cfg, w32).to(dt)
            with torch.inference_mode():
                y = model(latents.to(dt), t, text.to(dt), return_dict=False)[0]
            out["out_" + tag] = f32(y)
            # intermediate pins: rope table, condition embedder, first block
            with torch.inference_mode():
                rot = model.rope.forward_hidden_states(latents)
                out["rope_real"] = rot.real[0, 0].numpy()
                out["rope_imag"] = rot.imag[0, 0].numpy()
                temb, tproj, ctx, _, _ = model.condition_embedder(t, text.to(dt), None, None)
                out["temb_" + tag], out["tproj_" + tag], out["ctx_" + tag] = f32(temb), f32(tproj), f32(ctx)
                hs = model.patch_embedding(latents.to(dt)).flatten(2).transpose(1, 2)
                out["patch_" + tag] = f32(hs)
                b0 = model.blocks[0](hs.clone(), ctx, tproj.unflatten(1, (6, -1)), rot)
                out["block0_" + tag] = f32(b0)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})


def golden_attention():
    """The reference's own differential recipe (scripts/smoke_tests/test_attention_backends.py:232-388):
    B,H,S,D = 1,32,1024,128, seed 42, q,k,v = randn; gold = attention_register.get('sdpa').  On the CPU the
    recipe runs in fp32; we keep a 4-head slice of the gold output to stay small."""
    a = bootstrap.ref("src.attention.functions")
    torch.manual_seed(42)
    q = torch.randn(1, 32, 1024, 128)
    k = torch.randn(1, 32, 1024, 128)
    v = torch.randn(1, 32, 1024, 128)
    gold = a.attention_register.get("sdpa")(q, k, v)
    # cross-attention shaped case (Sq != Sk, ragged)
    torch.manual_seed(7)
    q2, k2, v2 = torch.randn(1, 2, 300, 128), torch.randn(1, 2, 77, 128), torch.randn(1, 2, 77, 128)
    gold2 = a.attention_register.get("sdpa")(q2, k2, v2)
    np.savez_compressed(os.path.join(GOLDEN, "attention.npz"), heads=np.array([0, 7, 19, 31]),
                        gold_1x32x1024x128_seed42=f32(gold[:, [0, 7, 19, 31]]),
                        q2=f32(q2), k2=f32(k2), v2=f32(v2), gold2=f32(gold2))
    print("attention", gold.shape)


def synthetic_model(sample: torch.Tensor, t: int) -> torch.Tensor:
    """Deterministic stand-in for the DiT inside scheduler goldens (same formula in the tests)."""
    return 0.35 * sample + 0.1 * torch.sin(sample * 3.0 + float(t) * 0.01)


def golden_scheduler():
    """Run the reference's own UniPCMultistepScheduler (scheduler/unipc.py) for several step counts."""
    u = bootstrap.ref("src.scheduler.unipc")
    out = {}
    for n, shift in ((4, 3.0), (8, 5.0), (50, 3.0), (50, 1.0)):
        sch = u.UniPCMultistepScheduler(shift=shift)
        sch.set_timesteps(n)
        tag = f"n{n}_s{shift:g}"
        out[tag + "_timesteps"] = sch.timesteps.numpy().astype(np.int64)
        out[tag + "_sigmas"] = sch.sigmas.numpy().astype(np.float32)
        x = torch.randn(1, 16, 2, 6, 8, generator=torch.Generator().manual_seed(42))
        out[tag + "_x0"] = f32(x)
        trace, norms = [], []
        for t in sch.timesteps:
            had_last = sch.last_sample is not None
            idx_before = sch.step_index
            mo = synthetic_model(x, int(t))
            x = sch.step(mo, t, x, return_dict=False)[0]
            idx = sch.step_index - 1
            use_corr = idx > 0 and (idx - 1) not in sch.disable_corrector and had_last
            trace.append((idx, sch.this_order, int(use_corr)))
            norms.append(float(x.double().norm()))
        out[tag + "_trace"] = np.array(trace, dtype=np.int64)
        out[tag + "_norms"] = np.array(norms, dtype=np.float64)
        out[tag + "_final"] = f32(x)
    np.savez_compressed(os.path.join(GOLDEN, "unipc.npz"), **out)
    print("scheduler", sorted(out)[:6], "...")


VAE_CFG = dict(base_dim=32, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2)
VAE_CASES = {
    # name: (latent shape, tiling, output subsample stride)
    "tiled": ((1, 16, 3, 36, 28), True, 3),      # tiles 32x28, 32x4, 12x28, 12x4 (stride 24) -> blends + crops
    "untiled": ((1, 16, 4, 20, 16), False, 2),   # one tile, 13 output frames
}


def golden_vae():
    """Reference AutoencoderKLWan (streaming, feat_cache) on a reduced-width decoder (base_dim 32) with the
    production tiling parameters (256/192 px).  Outputs are stored spatially subsampled to keep fixtures small."""
    import wan_vae

    v = bootstrap.ref("src.vae.wan.model")
    w = wan_vae.make_weights(base_dim=VAE_CFG["base_dim"], seed=7)
    out = {}
    for name, (shape, tiling, sub) in VAE_CASES.items():
        lat = torch.randn(shape, generator=torch.Generator().manual_seed(11))
        out[name + "_latents"] = f32(lat)
        for dt, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
            vae = v.AutoencoderKLWan(**VAE_CFG, temperal_downsample=[False, True, True]).eval()
            vae.load_state_dict(w, strict=False)
            vae = vae.to(dt)
            with torch.no_grad():
                z = vae.denormalize_latents(lat).to(dt)     # base_engine.py:2040-2048
                if tiling:
                    vae.enable_tiling()
                y = vae.decode(z, return_dict=False)[0]
            out[f"{name}_out_{tag}"] = f32(y[..., ::sub, ::sub])
            out[f"{name}_shape"] = np.array(y.shape)
    np.savez_compressed(os.path.join(GOLDEN, "wan_vae.npz"), **out)
    print("vae", {k: v.shape for k, v in out.items()})


FLUX_CONFIGS = {
    # name: (oracle make_weights kwargs, latent grid (h, w), text tokens)
    "flux_s56": (dict(dim=256, heads=2, num_layers=2, num_single_layers=2, in_channels=16, joint_dim=32, pooled_dim=24,
                      guidance_embeds=True), (6, 8), 8),
    "flux_s200": (dict(dim=256, heads=2, num_layers=1, num_single_layers=1, in_channels=16, joint_dim=32, pooled_dim=24,
                       guidance_embeds=False), (12, 16), 8),
}


def golden_flux():
    """Reference FluxTransformer2DModel (transformer/flux/base/model.py) with the `sdpa` backend on seeded inputs, fp32
    and bf16: final output plus pins after the embedders, the first dual-stream and the first single-stream block.
    timestep 0.5 and guidance 4.0 are exactly representable after the reference's bf16 `* 1000`."""
    import flux_dit

    fm = bootstrap.ref("src.transformer.flux.base.model")
    a = bootstrap.ref("src.attention.functions")
    a.attention_register.set_default("sdpa")
    for name, (cfg, (gh, gw), n_txt) in FLUX_CONFIGS.items():
        w32 = flux_dit.make_weights(**cfg, seed=1234, dtype=torch.float32)
        x = torch.randn(1, gh * gw, cfg["in_channels"], generator=torch.Generator().manual_seed(42))
        enc = torch.randn(1, n_txt, cfg["joint_dim"], generator=torch.Generator().manual_seed(43))
        pooled = torch.randn(1, cfg["pooled_dim"], generator=torch.Generator().manual_seed(44))
        t = torch.tensor([0.5])
        g = torch.tensor([4.0]) if cfg["guidance_embeds"] else None
        img_ids, txt_ids = flux_dit.latent_image_ids(gh, gw), torch.zeros(n_txt, 3)
        out = dict(hidden=f32(x), enc=f32(enc), pooled=f32(pooled), timestep=t.numpy(), img_ids=img_ids.numpy(),
                   txt_ids=txt_ids.numpy(), grid=np.array([gh, gw]))
        if g is not None:
            out["guidance"] = g.numpy()
        for dt, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
            model = fm.FluxTransformer2DModel(
                in_channels=cfg["in_channels"], num_layers=cfg["num_layers"], num_single_layers=cfg["num_single_layers"],
                attention_head_dim=cfg["dim"] // cfg["heads"], num_attention_heads=cfg["heads"],
                joint_attention_dim=cfg["joint_dim"], pooled_projection_dim=cfg["pooled_dim"],
                guidance_embeds=cfg["guidance_embeds"]).eval()
            model.load_state_dict(w32, strict=True)
            model = model.to(dt)
            with torch.inference_mode():
                y = model(x.to(dt), enc.to(dt), pooled.to(dt), t.to(dt), img_ids, txt_ids, g, return_dict=False)[0]
                out["out_" + tag] = f32(y)
                hs = model.x_embedder(x.to(dt))
                ts = t.to(dt) * 1000
                temb = (model.time_text_embed(ts, pooled.to(dt)) if g is None
                        else model.time_text_embed(ts, g.to(dt) * 1000, pooled.to(dt)))
                ctx = model.context_embedder(enc.to(dt))
                rope = model.pos_embed(torch.cat((txt_ids, img_ids), dim=0))
                out["temb_" + tag] = f32(temb)
                out["rope_cos"], out["rope_sin"] = rope[0].numpy(), rope[1].numpy()
                ctx1, hs1 = model.transformer_blocks[0](hidden_states=hs, encoder_hidden_states=ctx, temb=temb,
                                                        image_rotary_emb=rope)
                out["dual0_ctx_" + tag], out["dual0_x_" + tag] = f32(ctx1), f32(hs1)
                ctx2, hs2 = model.single_transformer_blocks[0](hidden_states=hs1, encoder_hidden_states=ctx1, temb=temb,
                                                               image_rotary_emb=rope)
                out["single0_ctx_" + tag], out["single0_x_" + tag] = f32(ctx2), f32(hs2)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})


HY15_CONFIGS = {
    # name: (oracle make_weights kwargs, latent shape, (text len, valid), (byt5 len, valid), image tokens, i2v?)
    "hy15_t2v": (dict(dim=256, heads=2, num_layers=2, num_refiner_layers=2, in_channels=9, out_channels=4, text_dim=48,
                      text2_dim=40, image_dim=24, byt5_hidden=64), (1, 9, 3, 4, 6), (10, 7), (6, 4), 5, False),
    "hy15_i2v": (dict(dim=256, heads=2, num_layers=1, num_refiner_layers=1, in_channels=9, out_channels=4, text_dim=48,
                      text2_dim=40, image_dim=24, byt5_hidden=64), (1, 9, 5, 8, 10), (12, 12), (6, 3), 5, True),
}


def build_reference_hy15(cfg, w32):
    hm = bootstrap.ref("src.transformer.hunyuanvideo15.base.model")
    model = hm.HunyuanVideo15Transformer3DModel(
        in_channels=cfg["in_channels"], out_channels=cfg["out_channels"], num_attention_heads=cfg["heads"],
        attention_head_dim=cfg["dim"] // cfg["heads"], num_layers=cfg["num_layers"],
        num_refiner_layers=cfg["num_refiner_layers"], text_embed_dim=cfg["text_dim"], text_embed_2_dim=cfg["text2_dim"],
        image_embed_dim=cfg["image_dim"], patch_size=1, patch_size_t=1).eval()
    # the reference hard-codes the ByT5 projection width to 2048 (model.py:856-858); the fixture keeps it small
    model.context_embedder_2 = hm.HunyuanVideo15ByT5TextProjection(cfg["text2_dim"], cfg["byt5_hidden"], cfg["dim"]).eval()
    model.load_state_dict(w32, strict=True)
    return model


def golden_hy15():
    """Reference HunyuanVideo15Transformer3DModel with the `sdpa` backend, fp32 and bf16: final output plus the reordered
    condition tokens and the first dual-stream block.  Text masks with padding (valid prefix) exercise the refiner's
    key-padding mask and the valid-first token reorder; t2v = all-zero image embeds, i2v = random image embeds."""
    import hy15_dit

    a = bootstrap.ref("src.attention.functions")
    a.attention_register.set_default("sdpa")
    for name, (cfg, lshape, (l1, v1), (l2, v2), l3, i2v) in HY15_CONFIGS.items():
        w32 = hy15_dit.make_weights(**cfg, seed=1234, dtype=torch.float32)
        x = torch.randn(lshape, generator=torch.Generator().manual_seed(42))
        text = torch.randn(1, l1, cfg["text_dim"], generator=torch.Generator().manual_seed(43))
        text2 = torch.randn(1, l2, cfg["text2_dim"], generator=torch.Generator().manual_seed(44))
        img = torch.randn(1, l3, cfg["image_dim"], generator=torch.Generator().manual_seed(45)) if i2v else torch.zeros(1, l3, cfg["image_dim"])
        m1, m2 = torch.zeros(1, l1), torch.zeros(1, l2)
        m1[:, :v1], m2[:, :v2] = 1, 1
        t = torch.tensor([500.0])
        out = dict(hidden=f32(x), text=f32(text), text2=f32(text2), image=f32(img), mask=m1.numpy(), mask2=m2.numpy(),
                   timestep=t.numpy())
        for dt, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
            model = build_reference_hy15(cfg, w32).to(dt)
            with torch.inference_mode():
                y = model(x.to(dt), t.to(dt), text.to(dt), m1, encoder_hidden_states_2=text2.to(dt),
                          encoder_attention_mask_2=m2, image_embeds=img.to(dt), return_dict=False)[0]
                out["out_" + tag] = f32(y)
                temb = model.time_embed(t.to(dt))
                out["temb_" + tag] = f32(temb)
                ref_txt = model.context_embedder(text.to(dt), t.to(dt), m1)
                out["refined_" + tag] = f32(ref_txt)
                rope = model.rope(x)
                out["rope_cos"], out["rope_sin"] = rope[0].numpy(), rope[1].numpy()
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})


QWEN_CONFIGS = {
    # name: (oracle make_weights kwargs, img_shapes of one sample [(f, h, w), ...], text tokens)
    "qwen_t2i": (dict(dim=256, heads=2, num_layers=2, in_channels=16, out_channels=4, joint_dim=48), [(1, 6, 8)], 9),
    "qwen_edit": (dict(dim=256, heads=2, num_layers=1, in_channels=16, out_channels=4, joint_dim=48),
                  [(1, 8, 10), (1, 6, 6), (1, 4, 10)], 13),          # noisy latent + two reference images (edit-plus)
}


def golden_qwen():
    """Reference QwenImageTransformer2DModel with the `sdpa` backend on seeded inputs, fp32 and bf16: output, time embedding,
    RoPE tables and the first block.  timestep 0.5 (sigma) is exact in bf16."""
    import qwen_dit

    qm = bootstrap.ref("src.transformer.qwenimage.base.model")
    a = bootstrap.ref("src.attention.functions")
    a.attention_register.set_default("sdpa")
    for name, (cfg, shapes, n_txt) in QWEN_CONFIGS.items():
        w32 = qwen_dit.make_weights(**cfg, seed=1234, dtype=torch.float32)
        n_img = sum(f * h * w_ for f, h, w_ in shapes)
        x = torch.randn(1, n_img, cfg["in_channels"], generator=torch.Generator().manual_seed(42))
        enc = torch.randn(1, n_txt, cfg["joint_dim"], generator=torch.Generator().manual_seed(43))
        t = torch.tensor([0.5])
        out = dict(hidden=f32(x), enc=f32(enc), timestep=t.numpy(), img_shapes=np.array(shapes))
        for dt, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
            model = qm.QwenImageTransformer2DModel(
                patch_size=2, in_channels=cfg["in_channels"], out_channels=cfg["out_channels"], num_layers=cfg["num_layers"],
                attention_head_dim=cfg["dim"] // cfg["heads"], num_attention_heads=cfg["heads"],
                joint_attention_dim=cfg["joint_dim"]).eval()
            model.load_state_dict(w32, strict=True)
            model = model.to(dt)
            with torch.inference_mode():
                y = model(hidden_states=x.to(dt), encoder_hidden_states=enc.to(dt), encoder_hidden_states_mask=torch.ones(1, n_txt),
                          timestep=t.to(dt), img_shapes=[shapes], txt_seq_lens=[n_txt], return_dict=False)[0]
                out["out_" + tag] = f32(y)
                hs = model.img_in(x.to(dt))
                temb = model.time_text_embed(t.to(dt), hs)
                out["temb_" + tag] = f32(temb)
                vf, tf = model.pos_embed([shapes], [n_txt], device=torch.device("cpu"))
                out["img_freqs_re"], out["img_freqs_im"] = vf.real.numpy(), vf.imag.numpy()
                out["txt_freqs_re"], out["txt_freqs_im"] = tf.real.numpy(), tf.imag.numpy()
                ctx = model.txt_in(model.txt_norm(enc.to(dt)))
                out["ctx_in_" + tag] = f32(ctx)
                c1, h1 = model.transformer_blocks[0](hidden_states=hs, encoder_hidden_states=ctx, encoder_hidden_states_mask=None,
                                                     temb=temb, image_rotary_emb=(vf, tf))
                out["block0_ctx_" + tag], out["block0_x_" + tag] = f32(c1), f32(h1)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})


HY15_VAE_CH = (128, 128, 64, 64, 32)          # decoder order; the production widths are (1024, 1024, 512, 256, 128)
HY15_VAE_CASES = {
    # name: (latent shape, tiling, spatial subsample stride of the stored output)
    "untiled": ((1, 32, 3, 8, 6), False, 1),
    "tiled": ((1, 32, 2, 14, 10), True, 2),       # 8x8-latent tiles at stride 6: 3 x 2 tiles, ragged last row / column
}


def golden_hy15vae():
    """Reference AutoencoderKLHunyuanVideo15 (reduced widths, production tiling parameters 128 px / overlap 0.25)."""
    import hy15_vae

    vm = bootstrap.ref("src.vae.hunyuanvideo15.model")
    w32 = hy15_vae.make_weights(HY15_VAE_CH, seed=7)
    out = {}
    for name, (shape, tiling, sub) in HY15_VAE_CASES.items():
        lat = torch.randn(shape, generator=torch.Generator().manual_seed(11))
        out[name + "_latents"] = f32(lat)
        for dt, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
            vae = vm.AutoencoderKLHunyuanVideo15(latent_channels=32, block_out_channels=tuple(reversed(HY15_VAE_CH))).eval()
            missing, unexpected = vae.load_state_dict(w32, strict=False)
            assert not unexpected and all(k.startswith("encoder.") for k in missing), (missing[:3], unexpected[:3])
            vae = vae.to(dt)
            if tiling:
                vae.enable_tiling()
            with torch.no_grad():
                y = vae.decode(lat.to(dt), return_dict=False)[0]
            out[f"{name}_out_{tag}"] = f32(y[..., ::sub, ::sub])
            out[f"{name}_shape"] = np.array(y.shape)
    np.savez_compressed(os.path.join(GOLDEN, "hy15_vae.npz"), **out)
    print("hy15_vae", {k: v.shape for k, v in out.items()})


FLUX2_CONFIGS = {
    # name: (oracle make_weights kwargs, latent grid (h, w), text tokens)
    "flux2_s56": (dict(dim=256, heads=2, num_layers=2, num_single_layers=2, in_channels=16, joint_dim=48, guidance_embeds=False), (6, 8), 8),
    "flux2_s200": (dict(dim=256, heads=2, num_layers=1, num_single_layers=1, in_channels=16, joint_dim=48, guidance_embeds=True), (12, 16), 8),
}


def flux2_ids(gh, gw, n_txt):
    """4-axis position ids as the Flux2 engine prepares them: image (t=0, h, w, l=0), text (0, 0, 0, l)."""
    img = torch.zeros(gh * gw, 4)
    img[:, 1] = torch.arange(gh * gw) // gw
    img[:, 2] = torch.arange(gh * gw) % gw
    txt = torch.zeros(n_txt, 4)
    txt[:, 3] = torch.arange(n_txt)
    return img, txt


def golden_flux2():
    """Reference Flux2Transformer2DModel (BASELINE configs[0] family) with the `sdpa` backend, fp32 and bf16."""
    import flux2_dit

    fm = bootstrap.ref("src.transformer.flux2.base.model")
    a = bootstrap.ref("src.attention.functions")
    a.attention_register.set_default("sdpa")
    for name, (cfg, (gh, gw), n_txt) in FLUX2_CONFIGS.items():
        w32 = flux2_dit.make_weights(**cfg, seed=1234, dtype=torch.float32)
        x = torch.randn(1, gh * gw, cfg["in_channels"], generator=torch.Generator().manual_seed(42))
        enc = torch.randn(1, n_txt, cfg["joint_dim"], generator=torch.Generator().manual_seed(43))
        t = torch.tensor([0.5])
        g = torch.tensor([4.0]) if cfg["guidance_embeds"] else None
        img_ids, txt_ids = flux2_ids(gh, gw, n_txt)
        out = dict(hidden=f32(x), enc=f32(enc), timestep=t.numpy(), img_ids=img_ids.numpy(), txt_ids=txt_ids.numpy())
        if g is not None:
            out["guidance"] = g.numpy()
        for dt, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
            model = fm.Flux2Transformer2DModel(
                in_channels=cfg["in_channels"], num_layers=cfg["num_layers"], num_single_layers=cfg["num_single_layers"],
                attention_head_dim=cfg["dim"] // cfg["heads"], num_attention_heads=cfg["heads"],
                joint_attention_dim=cfg["joint_dim"], guidance_embeds=cfg["guidance_embeds"]).eval()
            model.load_state_dict(w32, strict=True)
            model = model.to(dt)
            with torch.inference_mode():
                y = model(x.to(dt), enc.to(dt), t.to(dt), img_ids, txt_ids, None if g is None else g.to(dt), return_dict=False)[0]
            out["out_" + tag] = f32(y)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    which = sys.argv[1:] or ["dit", "attention", "scheduler", "vae", "flux", "hy15", "qwen", "hy15vae", "flux2"]
    for wname in which:
        fn = globals().get("golden_" + wname)
        if fn is None:
            print("skip", wname)
            continue
        fn()
