"""GPU parity at the BASELINE widths (VERDICT r1 "Next round" item 1): the kernels run together in the shapes the
headline benchmark uses -- GROUP_M rasterisation, the CTA-pair GEMM, 40-head attention grids, 5120-wide row kernels,
the 96-channel (BK = 32) conv path -- and are compared with the oracle on **bf16-rounded weights and inputs**, so
that weight quantisation (which made "ours" and "reference bf16" read the same 6.2e-3 on the toy goldens) is not
part of either error.

Three numbers per case:
  exact   = oracle in fp32 math on the bf16-rounded weights/inputs (no interior rounding at all);
  ref16   = the oracle's bf16 path = the reference's own bf16 pipeline (bit-exact restatement, tests/test_oracle_*);
  ours    = this repo's CUDA path through the C ABI.
Bars (stated below): rel-L2(ours, exact) <= BAR and <= rel-L2(ref16, exact) -- the CUDA path must be CLOSER to exact
math than the reference's own bf16 pipeline is.  Why the bars are not 1e-3: every tensor-core operand is a bf16 tensor
in HBM (the reference's too), and ONE bf16 rounding of a general-position tensor is 2^-7.5 / sqrt(12) = 1.6e-3 relative
L2.  A block chains >= 4 such roundings per sub-layer (modulated-norm output, q|k|v, attention output / FFN hidden,
residual store) x 3 sub-layers; with independent errors that is a floor of ~4e-3 for ANY bf16-operand pipeline, and
the 35-conv VAE decoder chains ~35.  Measured on the B200 (profiles/r02_baseline_width_parity.jsonl): block 4.37e-3
(reference pipeline 5.07e-3), block update 5.21e-3 (6.04e-3), VAE tile 9.93e-3 (1.40e-2).  The errors are printed
(pytest -s) and appended to gpurun_out/baseline_width_parity.jsonl when that directory exists.
"""
import json
import os
import time

import pytest
import torch

import wan_dit
import wan_vae

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _record(name, **vals):
    print(f"[baseline-width] {name}: " + ", ".join(f"{k}={v:.3e}" if isinstance(v, float) else f"{k}={v}"
                                                     for k, v in vals.items()))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "baseline_width_parity.jsonl"), "a") as f:
            f.write(json.dumps(dict(case=name, **vals)) + "\n")


# Stated bars.  DiT block: h_out is the bf16 residual stream after three read-modify-write stores.
BAR_BLOCK = 4.8e-3             # measured 4.37e-3 (reference bf16 pipeline: 5.07e-3)
BAR_BLOCK_UPDATE = 5.7e-3      # error of (h_out - h_in) relative to the update's norm; measured 5.21e-3 (reference 6.04e-3)
BAR_VAE_TILE = 1.1e-2          # measured 9.93e-3 (reference bf16 pipeline: 1.40e-2)


def test_wan_block_at_a14b_width_vs_oracle():
    """One WanTransformerBlock (model.py:1101-1333) at d = 5120 / 40 heads / ffn 13,824 / 512 text tokens on the RoPE
    grid (4, 32, 32) -> S = 4096 tokens (32 query tiles x 40 heads; 16 pair-tiles of rows in every GEMM)."""
    from apex_studio_b200.wan import WanConfig, WanTransformer3DModel
    from apex_studio_b200.wan.model import _Workspace

    dim, heads, ffn, L, grid = 5120, 40, 13824, 512, (4, 32, 32)
    S = grid[0] * grid[1] * grid[2]
    w32 = wan_dit.make_weights(dim=dim, heads=heads, ffn_dim=ffn, num_layers=1, text_dim=64, freq_dim=256, seed=77)
    w16 = {k: v.bfloat16() for k, v in w32.items()}                  # the blanket module.to(bf16) of to_mixin.py:358
    wq = {k: v.float() for k, v in w16.items()}                      # the same values, held in fp32 for exact math
    g = torch.Generator().manual_seed(78)
    h0 = torch.randn(1, S, dim, generator=g).bfloat16()
    ctx = torch.randn(1, L, dim, generator=g).bfloat16()
    temb6 = (torch.randn(1, 6, dim, generator=g) * 0.5).bfloat16()   # O(1) gates / scales so the updates matter
    freqs = wan_dit.rope_table(128, grid)

    t0 = time.time()
    exact = wan_dit.block_forward(h0.float(), ctx.float(), temb6.float(), freqs, wq, "blocks.0", heads)
    t1 = time.time()
    ref16 = wan_dit.block_forward(h0, ctx, temb6, freqs, w16, "blocks.0", heads)
    t2 = time.time()

    model = WanTransformer3DModel(WanConfig(num_attention_heads=heads, ffn_dim=ffn, num_layers=1, text_dim=64))
    model.load_state_dict(w32, device=DEV)
    h = h0[0].to(DEV).clone()
    ws = _Workspace(S, L, model.config, torch.device(DEV))
    model.text_kv(ctx[0].to(DEV), ws)            # to_k | to_v + norm_k of the cross-attention, hoisted out of the block
    model.block(0, h, ctx[0].to(DEV), temb6[0].to(DEV), model._rope(grid), ws)
    torch.cuda.synchronize()

    ours, ref_err = rel_l2(h[None], exact), rel_l2(ref16, exact)
    upd = exact - h0.float()
    ours_u = rel_l2(h[None].float().cpu() - h0.float(), upd)
    ref_u = rel_l2(ref16.float() - h0.float(), upd)
    _record("wan_block_d5120_h40_ffn13824_S4096", ours=ours, ref_bf16=ref_err, ours_update=ours_u, ref_bf16_update=ref_u,
            ours_vs_ref_bf16=rel_l2(h[None], ref16), update_over_h=(upd.norm() / h0.float().norm()).item(),
            oracle_fp32_s=t1 - t0, oracle_bf16_s=t2 - t1)
    assert torch.isfinite(h).all()
    assert ours <= BAR_BLOCK and ours <= ref_err, (ours, ref_err)
    assert ours_u <= BAR_BLOCK_UPDATE and ours_u <= ref_u, (ours_u, ref_u)


def test_wan_vae_tile_at_base_dim_96_vs_oracle():
    """One decoder tile (vae/wan/model.py:972-1021) at the production widths 384/384/384/192/96: 16 x 16 latents x 3
    latent frames -> 128 x 128 px x 9 frames (conv BK = 32 path for Cin = 96, 192-wide N tiles, mid attention at C = 384)."""
    from apex_studio_b200.vae import AutoencoderKLWan, WanVAEConfig

    w32 = wan_vae.make_weights(base_dim=96, seed=11)
    w16 = {k: v.bfloat16() for k, v in w32.items()}
    wq = {k: v.float() for k, v in w16.items()}
    g = torch.Generator().manual_seed(12)
    z = torch.randn(1, 16, 3, 16, 16, generator=g).bfloat16()
    t0 = time.time()
    exact = wan_vae.decoder_forward(z.float(), wq)
    t1 = time.time()
    ref16 = wan_vae.decoder_forward(z, w16)
    t2 = time.time()
    vae = AutoencoderKLWan(WanVAEConfig(base_dim=96))
    vae.load_state_dict(w32, device=DEV)
    out = vae.decode_tile(z[0].to(DEV))
    torch.cuda.synchronize()
    assert tuple(out.shape) == tuple(exact.shape[1:])
    ours, ref_err = rel_l2(out[None], exact), rel_l2(ref16, exact)
    _record("wan_vae_tile_base96_16x16x3", ours=ours, ref_bf16=ref_err, ours_vs_ref_bf16=rel_l2(out[None], ref16),
            oracle_fp32_s=t1 - t0, oracle_bf16_s=t2 - t1)
    assert ours <= BAR_VAE_TILE and ours <= ref_err, (ours, ref_err)


@pytest.mark.parametrize("epi", ["gelu", "gate_res"])
def test_linear_full_height_epilogues(epi):
    """M = 75,600 (591 row tiles, ragged last pair) with the GELU-tanh and gate+residual epilogues -- the two that run at
    full height in every block (ffn.net.0.proj, to_out / ffn.net.2); round 1 skipped them."""
    from apex_studio_b200 import ops

    M, N, K = 75600, 512, 640
    torch.manual_seed(5)
    x = torch.randn(M, K, device=DEV).bfloat16()
    w = (torch.randn(N, K, device=DEV) * 0.05).bfloat16()
    b = (torch.randn(N, device=DEV) * 0.5).bfloat16()
    acc = x.float() @ w.float().t() + b.float()
    if epi == "gelu":
        ref, out = torch.nn.functional.gelu(acc, approximate="tanh"), ops.linear(x, w, b, epilogue=ops.EPI_GELU_TANH)
    else:
        h = torch.randn(M, N, device=DEV).bfloat16()
        gt = torch.randn(N, device=DEV).bfloat16()
        ref = h.float() + gt.float() * acc
        out = h.clone()
        ops.linear(x, w, b, epilogue=ops.EPI_GATE_RES, out=out, gate=gt)
    err = rel_l2(out, ref)
    _record(f"linear_M75600_{epi}", ours=err)
    assert err <= 4e-3
    # last (ragged) row tile and first tile both written
    assert rel_l2(out[-200:], ref[-200:]) <= 4e-3 and rel_l2(out[:200], ref[:200]) <= 4e-3
