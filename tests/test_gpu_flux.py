"""GPU parity tests of the Flux path (BASELINE.json configs[1]): every call goes through the C ABI
and is checked against the CPU oracle (oracle/flux_dit.py, pinned bit-exactly to the reference's own FluxTransformer2DModel)
and the golden vectors recorded from the reference (tests/golden/flux_*.npz).

Tolerances:
  * row kernels (per-head RMS-norm + rotation, AdaLayerNormZero modulation, SwiGLU) reproduce the reference's bf16 rounding
    points: bit-exact, or <= 1 bf16 ulp (on the row scale) on <= 0.2 % of elements where the fp32 reduction order of the row
    statistic / the fp32 intrinsic flips a rounding;
  * whole forward / blocks: relative L2 vs the exact-math fp32 oracle <= max(1e-3, 1.5 x the reference's own bf16 error
    against that oracle), and relative L2 <= 2e-2 vs the reference's bf16 golden output;
  * full FLUX.1-dev width (d = 3072, 24 heads, 4096 image + 512 text tokens) on one dual + one single block vs the exact
    oracle: same bar."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import flux_dit
from conftest import GOLDEN
from test_gpu_parity import _ulp_report, rel_l2
from test_oracle_flux import CONFIGS, inputs, kw, load

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from apex_studio_b200 import ops as _ops

    return _ops


# ------------------------------------------------------------------------------------------------ row kernels
def _norm_ref(x, w, mode, eps=1e-6):
    """x [rows, heads, 128] bf16 on the CPU, rounding points of the three families (csrc/mmdit_ops.cu header)."""
    if mode == 0:
        return x
    if mode == 1:    # torch.nn.RMSNorm (flux/base/model.py:102-107)
        return F.rms_norm(x, (128,), w, eps)
    r = x.float().pow(2).mean(-1, keepdim=True).add(eps).rsqrt()
    if mode == 2:    # InplaceRMSNorm (efficiency/mod.py:24-35)
        y = x * r.to(x.dtype)
        return y * w.to(x.dtype)
    y = (x * r).to(w.dtype)   # diffusers RMSNorm (bf16 * fp32 promotes to fp32, then the cast to the weight dtype)
    return y * w


def _rope_ref(x, cos, sin):
    """x [rows, heads, 128] bf16; cos/sin fp32 [rows, 64]: fp32 math, one rounding (diffusers apply_rotary_emb)."""
    c, s = cos.repeat_interleave(2, dim=1)[:, None, :], sin.repeat_interleave(2, dim=1)[:, None, :]
    re, im = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-im, re], dim=-1).flatten(2)
    return (x.float() * c + rot.float() * s).to(x.dtype)


@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("rows,heads", [(56, 2), (777, 24)])
def test_headnorm_rope_vs_reference_rounding(ops, mode, rows, heads):
    torch.manual_seed(rows + mode)
    d = heads * 128
    buf = (torch.randn(rows, 3 * d) * 2).bfloat16()
    wq, wk = (1 + 0.1 * torch.randn(128)).bfloat16(), (1 + 0.1 * torch.randn(128)).bfloat16()
    ang = torch.rand(rows, 64, dtype=torch.float64) * 6.28
    cos, sin = ang.cos().float(), ang.sin().float()
    if mode == 2:   # HunyuanVideo-1.5 casts the table to the activation dtype first (efficiency/ops.py:201-202)
        cos, sin = cos.bfloat16().float(), sin.bfloat16().float()
    table = torch.stack([cos, sin], dim=-1).contiguous()
    q, k = buf[:, :d].reshape(rows, heads, 128), buf[:, d:2 * d].reshape(rows, heads, 128)
    ref_q = _rope_ref(_norm_ref(q, wq, mode), cos, sin).reshape(rows, d)
    ref_k = _rope_ref(_norm_ref(k, wk, mode), cos, sin).reshape(rows, d)
    g = buf.to(DEV)
    ops.headnorm_rope_(g[:, :d], g[:, d:2 * d], wq.to(DEV), wk.to(DEV), table.to(DEV), heads, 1e-6, mode)
    out = g.cpu()
    assert torch.equal(out[:, 2 * d:], buf[:, 2 * d:])                       # v untouched
    for got, ref in ((out[:, :d], ref_q), (out[:, d:2 * d], ref_k)):
        frac, ulps = _ulp_report(got, ref)
        assert frac <= 2e-3 and ulps <= 1.01, (mode, frac, ulps)
    # norm only, single tensor, no gain
    g2 = buf.to(DEV)
    ops.headnorm_rope_(g2[:, :d], None, None, None, None, heads, 1e-6, mode)
    frac, ulps = _ulp_report(g2[:, :d].cpu(), _norm_ref(q, torch.ones(128).bfloat16(), mode).reshape(rows, d))
    assert frac <= 2e-3 and ulps <= 1.01, (mode, frac, ulps)
    assert torch.equal(g2[:, d:].cpu(), buf[:, d:])
    # rotation only is exact up to the fp32 intrinsic order: identity table leaves the data untouched
    g3 = buf.to(DEV)
    ident = torch.stack([torch.ones(rows, 64), torch.zeros(rows, 64)], dim=-1).contiguous().to(DEV)
    ops.headnorm_rope_(g3[:, :d], g3[:, d:2 * d], None, None, ident, heads, 1e-6, 0)
    assert torch.equal(g3.cpu(), buf)


def test_headnorm_rope_argument_errors(ops):
    x = torch.zeros(8, 256, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        ops.headnorm_rope_(x[:, :100], None, None, None, None, 2)
    with pytest.raises(ValueError):
        ops.headnorm_rope_(x, None, None, None, torch.zeros(8, 64, 2, device=DEV, dtype=torch.bfloat16), 2)
    with pytest.raises(ValueError):
        ops.headnorm_rope_(x, None, None, None, None, 2, norm_mode=7)
    with pytest.raises(ValueError):
        ops.headnorm_rope_(x.cpu(), None, None, None, None, 2)


@pytest.mark.parametrize("rows,dim", [(5, 256), (333, 3072), (64, 2048)])
def test_adaln_zero_modulate_vs_oracle(ops, rows, dim):
    torch.manual_seed(rows + dim)
    x = (torch.randn(rows, dim) * 3 + 0.5).bfloat16()
    scale, shift = (0.3 * torch.randn(1, dim)).bfloat16(), (0.3 * torch.randn(1, dim)).bfloat16()
    ref = flux_dit.ada_modulate(x[None], scale, shift)[0]
    wide = torch.zeros(rows, 2 * dim, device=DEV, dtype=torch.bfloat16)      # strided output rows
    out = ops.adaln_zero_modulate(x.to(DEV), scale[0].to(DEV), shift[0].to(DEV), out=wide[:, dim:]).cpu()
    frac, ulps = _ulp_report(out, ref)
    assert frac <= 2e-3 and ulps <= 1.01, (frac, ulps)
    assert wide[:, :dim].abs().max().item() == 0


def test_swiglu_vs_torch(ops):
    torch.manual_seed(3)
    x = (torch.randn(300, 2 * 1024) * 2).bfloat16()
    ref = F.silu(x[:, :1024]) * x[:, 1024:]
    out = ops.swiglu(x.to(DEV)).cpu()
    frac, ulps = _ulp_report(out, ref)
    assert frac <= 2e-3 and ulps <= 1.01, (frac, ulps)


# ------------------------------------------------------------------------------------------------ model
def _model(cfg, w32):
    from apex_studio_b200.flux import FluxConfig, FluxTransformer2DModel

    m = FluxTransformer2DModel(FluxConfig(
        in_channels=cfg["in_channels"], num_layers=cfg["num_layers"], num_single_layers=cfg["num_single_layers"],
        num_attention_heads=cfg["heads"], joint_attention_dim=cfg["joint_dim"], pooled_projection_dim=cfg["pooled_dim"],
        guidance_embeds=cfg["guidance_embeds"]))
    m.load_state_dict(w32, device=DEV)
    return m


@pytest.mark.parametrize("name", list(CONFIGS))
def test_flux_forward_vs_reference_golden(name):
    cfg, g = CONFIGS[name], load(name)
    w32 = flux_dit.make_weights(**cfg, seed=1234, dtype=torch.float32)
    m = _model(cfg, w32)
    x, enc, pooled, t, img_ids, txt_ids, guidance = inputs(g, torch.bfloat16)
    out = m(x.to(DEV), enc.to(DEV), pooled.to(DEV), t.to(DEV), img_ids, txt_ids, guidance, return_dict=False)[0]
    assert out.shape == x.shape and out.dtype == torch.bfloat16
    exact, ref16 = torch.from_numpy(g["out_fp32"]), torch.from_numpy(g["out_bf16"])
    ours, theirs = rel_l2(out, exact), rel_l2(ref16, exact)
    assert ours <= max(1e-3, 1.5 * theirs), (ours, theirs)
    assert rel_l2(out, ref16) <= 2e-2
    # residual stream after the whole forward is finite and the workspace is reused on a second call (same bits)
    out2 = m(x.to(DEV), enc.to(DEV), pooled.to(DEV), t.to(DEV), img_ids, txt_ids, guidance, return_dict=False)[0]
    assert torch.equal(out, out2)


def test_flux_blocks_vs_reference_golden():
    """First dual-stream and first single-stream block in isolation (model.py:257-328, :194-228)."""
    import torch.nn.functional as F2
    from apex_studio_b200 import ops
    from apex_studio_b200.flux.model import _Workspace

    name = "flux_s200"
    cfg, g = CONFIGS[name], load(name)
    w32 = flux_dit.make_weights(**cfg, seed=1234, dtype=torch.float32)
    m = _model(cfg, w32)
    x, enc, pooled, t, img_ids, txt_ids, guidance = inputs(g, torch.bfloat16)
    n_txt, n_img = enc.shape[1], x.shape[1]
    temb = torch.from_numpy(g["temb_bf16"]).bfloat16().to(DEV)
    got_temb = m.time_text_embed(t.to(DEV) * 1000, None, pooled.to(DEV))
    assert rel_l2(got_temb, torch.from_numpy(g["temb_fp32"])) <= 1e-2
    mod_all = ops.linear(F2.silu(temb), m.w["modulation.weight"], m.w["modulation.bias"])[0]
    rope = m._rope(txt_ids, img_ids)
    ws = _Workspace(n_txt, n_img, m.config, DEV)
    ops.linear(enc[0].to(DEV), m.w["context_embedder.weight"], m.w["context_embedder.bias"], out=ws.h[:n_txt])
    ops.linear(x[0].to(DEV), m.w["x_embedder.weight"], m.w["x_embedder.bias"], out=ws.h[n_txt:])
    m.dual_block(0, ws, mod_all, rope)
    for rows, key in ((slice(n_txt, None), "dual0_x"), (slice(0, n_txt), "dual0_ctx")):
        exact, ref16 = torch.from_numpy(g[key + "_fp32"])[0], torch.from_numpy(g[key + "_bf16"])[0]
        ours, theirs = rel_l2(ws.h[rows], exact), rel_l2(ref16, exact)
        assert ours <= max(1e-3, 1.5 * theirs), (key, ours, theirs)
    m.single_block(0, ws, mod_all, rope)
    for rows, key in ((slice(n_txt, None), "single0_x"), (slice(0, n_txt), "single0_ctx")):
        exact, ref16 = torch.from_numpy(g[key + "_fp32"])[0], torch.from_numpy(g[key + "_bf16"])[0]
        ours, theirs = rel_l2(ws.h[rows], exact), rel_l2(ref16, exact)
        assert ours <= max(1e-3, 2.0 * theirs), (key, ours, theirs)   # errors of two blocks compound


def test_flux_dev_width_one_dual_one_single_block_vs_exact_oracle():
    """FLUX.1-dev shapes (d = 3072, 24 x 128 heads, 64x64 latent grid = 4096 image tokens + 512 text tokens, guidance
    embedding) with 1 + 1 layers: CUDA path vs the exact-math fp32 oracle, bar = 1.5 x the bf16 error of the oracle run in
    bf16 (= the reference's arithmetic)."""
    cfg = dict(dim=3072, heads=24, num_layers=1, num_single_layers=1, in_channels=64, joint_dim=4096, pooled_dim=768,
               guidance_embeds=True)
    w32 = flux_dit.make_weights(**cfg, seed=5, dtype=torch.float32, std=0.02)
    m = _model(cfg, w32)
    gen = torch.Generator().manual_seed(42)
    x, enc, pooled = torch.randn(1, 4096, 64, generator=gen), torch.randn(1, 512, 4096, generator=gen), torch.randn(1, 768, generator=gen)
    t, guidance = torch.tensor([0.5]), torch.tensor([4.0])
    img_ids, txt_ids = flux_dit.latent_image_ids(64, 64), torch.zeros(512, 3)
    out = m(x.to(DEV), enc.to(DEV), pooled.to(DEV), t.to(DEV), img_ids, txt_ids, guidance.to(DEV), return_dict=False)[0]
    k = kw(cfg)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    exact = flux_dit.flux_forward(x, enc, pooled, t, img_ids, txt_ids, guidance, w32, **k)
    bf = flux_dit.flux_forward(x.bfloat16(), enc.bfloat16(), pooled.bfloat16(), t.bfloat16(), img_ids, txt_ids, guidance,
                               {kk: v.bfloat16() for kk, v in w32.items()}, **k)
    ours, theirs = rel_l2(out, exact), rel_l2(bf, exact)
    assert torch.isfinite(out).all()
    assert ours <= max(1e-3, 1.5 * theirs), (ours, theirs)


def test_flux_denoise_loop_vs_oracle():
    """engine/flux/shared.py:504-620 for 3 steps on the small golden model: product loop (CUDA forward + flow-match Euler)
    vs the oracle loop run in bf16 on the CPU."""
    from apex_studio_b200.denoise import flux_denoise
    from apex_studio_b200.scheduler import FlowMatchEulerDiscreteScheduler, calculate_shift

    name = "flux_s56"
    cfg, g = CONFIGS[name], load(name)
    w32 = flux_dit.make_weights(**cfg, seed=1234, dtype=torch.float32)
    m = _model(cfg, w32)
    x, enc, pooled, _, img_ids, txt_ids, guidance = inputs(g, torch.bfloat16)
    n = 3
    sch = FlowMatchEulerDiscreteScheduler()
    ts = sch.set_timesteps(n, device=DEV, sigmas=np.linspace(1.0, 1 / n, n), mu=calculate_shift(x.shape[1]))
    out = flux_denoise(latents=x.to(DEV), timesteps=ts, scheduler=sch, transformer=m, prompt_embeds=enc.to(DEV),
                       pooled_prompt_embeds=pooled.to(DEV), latent_ids=img_ids, text_ids=txt_ids, guidance=guidance.to(DEV))
    wb = {k_: v.bfloat16() for k_, v in w32.items()}
    ref = flux_dit.denoise(x, enc, pooled, img_ids, txt_ids, guidance, wb, n, **kw(cfg))
    exact = flux_dit.denoise(x.float(), enc.float(), pooled.float(), img_ids, txt_ids, guidance, w32, n, **kw(cfg))
    assert out.dtype == torch.bfloat16 and sch.step_index == n
    ours, theirs = rel_l2(out, exact), rel_l2(ref, exact)
    assert ours <= max(1e-3, 1.5 * theirs), (ours, theirs)
    assert rel_l2(out, ref) <= 2e-2, rel_l2(out, ref)


def test_cuda_graph_replay_is_bit_identical_to_eager():
    """one whole Flux forward captured into a CUDA graph (apex_studio_b200.graph.GraphedCallable) and replayed
    with new latents gives the eager result bit for bit."""
    from apex_studio_b200.graph import GraphedCallable

    name = "flux_s200"
    cfg, g = CONFIGS[name], load(name)
    m = _model(cfg, flux_dit.make_weights(**cfg, seed=1234, dtype=torch.float32))
    x, enc, pooled, t, img_ids, txt_ids, guidance = inputs(g, torch.bfloat16)
    x, enc, pooled, t = x.to(DEV), enc.to(DEV), pooled.to(DEV), t.to(DEV)
    fwd = lambda xx, ee, pp, tt: m(xx, ee, pp, tt, img_ids, txt_ids, None, return_dict=False)[0]
    graphed = GraphedCallable(fwd, (x, enc, pooled, t))
    x2 = (x.float() * 0.5 + 0.1).bfloat16()
    for xi in (x, x2, x):
        eager = fwd(xi, enc, pooled, t).clone()
        assert torch.equal(graphed(xi, enc, pooled, t), eager)
    with pytest.raises(ValueError):
        graphed(x[:, :10], enc, pooled, t)
