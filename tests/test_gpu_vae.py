"""GPU parity tests of the Wan VAE decode path (C ABI: b200_conv3d_cl, b200_rmsnorm_silu_cl, b200_upsample2x_cl,
b200_softmax_rows, b200_blend_tile, b200_linear) against oracle/wan_vae.py and the golden vectors recorded from the
reference's AutoencoderKLWan.

Tolerances: convolutions -- relative L2 <= 4e-3 vs fp32 math on the same bf16 operands (one bf16 output rounding);
copy-like kernels (upsample, blend) -- exact; whole decode -- relative L2 vs the fp32 oracle <= max(1e-2, 1.5 x the
reference's own bf16 error vs that oracle) and relative L2 <= 3e-2 vs the reference's bf16 output.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import wan_vae
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _cl(x):  # [1,C,T,H,W] -> channels-last [T,H,W,C] bf16 on the GPU
    return x[0].permute(1, 2, 3, 0).contiguous().to(DEV, torch.bfloat16)


def _tap_major(w):
    from apex_studio_b200.vae.wan import AutoencoderKLWan

    return AutoencoderKLWan._tap_major(w).to(DEV, torch.bfloat16)


@pytest.mark.parametrize("T,H,W,cin,cout,taps", [
    (3, 8, 8, 32, 32, (3, 3, 3)),       # BK=32 path (SWIZZLE_64B), tiny
    (2, 12, 28, 64, 96, (3, 3, 3)),     # ragged W (28 -> BW 16), BK=64
    (5, 18, 16, 96, 96, (3, 3, 3)),     # Cin 96 = 3 chunks of 32
    (4, 9, 5, 128, 384, (3, 3, 3)),     # tiny W, two N tiles of 192
    (3, 32, 32, 384, 384, (3, 3, 3)),   # production first-stage shape
    (2, 40, 136, 192, 96, (1, 3, 3)),   # per-frame Conv2d of WanResample, W > 128
    (6, 6, 10, 64, 128, (3, 1, 1)),     # time_conv taps
])
def test_conv3d_vs_torch(T, H, W, cin, cout, taps):
    from apex_studio_b200.vae.wan import conv3d_cl

    g = torch.Generator().manual_seed(T * 100 + H + W + cin)
    x = torch.randn(1, cin, T, H, W, generator=g).bfloat16()
    w = (torch.randn(cout, cin, *taps, generator=g) * (cin * taps[0] * taps[1] * taps[2]) ** -0.5).bfloat16()
    b = (torch.randn(cout, generator=g) * 0.1).bfloat16()
    ref = wan_vae.causal_conv3d(x.float(), {"c.weight": w.float(), "c.bias": b.float()}, "c")   # fp32 math
    out = conv3d_cl(_cl(x), _tap_major(w.float()), b.to(DEV), taps, cout)
    got = out.permute(3, 0, 1, 2).unsqueeze(0)
    assert rel_l2(got, ref) <= 4e-3
    # residual epilogue
    res = torch.randn(1, cout, T, H, W, generator=g).bfloat16()
    out2 = conv3d_cl(_cl(x), _tap_major(w.float()), b.to(DEV), taps, cout, residual=_cl(res))
    assert rel_l2(out2.permute(3, 0, 1, 2).unsqueeze(0), ref + res.float()) <= 4e-3


def test_conv3d_planar_and_interleave_outputs():
    from apex_studio_b200.vae.wan import conv3d_cl

    g = torch.Generator().manual_seed(5)
    T, H, W, C = 4, 10, 12, 64
    x = torch.randn(1, C, T, H, W, generator=g).bfloat16()
    # conv_out: 3 real channels padded to 16, planar output
    w = (torch.randn(3, C, 3, 3, 3, generator=g) * (27 * C) ** -0.5).bfloat16()
    b = (torch.randn(3, generator=g) * 0.1).bfloat16()
    ref = wan_vae.causal_conv3d(x.float(), {"c.weight": w.float(), "c.bias": b.float()}, "c")[0]
    wp = F.pad(w.float(), (0, 0, 0, 0, 0, 0, 0, 0, 0, 13))
    out = conv3d_cl(_cl(x), _tap_major(wp), F.pad(b, (0, 13)).to(DEV), (3, 3, 3), 16, planar_channels=3)
    assert out.shape == (3, T, H, W) and rel_l2(out, ref) <= 4e-3
    # time_conv + 2x temporal interleave (upsample3d): frame 0 bypasses, frames 1.. doubled
    wt = (torch.randn(2 * C, C, 3, 1, 1, generator=g) * (3 * C) ** -0.5).bfloat16()
    bt = (torch.randn(2 * C, generator=g) * 0.1).bfloat16()
    rest = wan_vae.causal_conv3d(x[:, :, 1:].float(), {"c.weight": wt.float(), "c.bias": bt.float()}, "c")
    rest = rest.reshape(1, 2, C, T - 1, H, W)
    expect = torch.cat([x[:, :, :1].float(), torch.stack((rest[:, 0], rest[:, 1]), 3).reshape(1, C, 2 * (T - 1), H, W)], 2)
    xcl = _cl(x)
    y = torch.empty(1 + 2 * (T - 1), H, W, C, device=DEV, dtype=torch.bfloat16)
    y[0].copy_(xcl[0])
    conv3d_cl(xcl[1:], _tap_major(wt.float()), bt.to(DEV), (3, 1, 1), 2 * C, out=y, interleave=True)
    assert rel_l2(y.permute(3, 0, 1, 2).unsqueeze(0), expect) <= 4e-3


@pytest.mark.parametrize("C", [32, 96, 192, 384])
def test_rmsnorm_silu_and_upsample(C):
    from apex_studio_b200.vae.wan import rmsnorm_silu_cl, upsample2x_cl

    g = torch.Generator().manual_seed(C)
    x = torch.randn(1, C, 3, 7, 9, generator=g).bfloat16()
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).bfloat16()
    ref = F.silu(wan_vae.rms_norm(x.float(), gamma.float().view(C, 1, 1, 1)))
    out = rmsnorm_silu_cl(_cl(x), gamma.to(DEV)).permute(3, 0, 1, 2).unsqueeze(0)
    assert rel_l2(out, ref) <= 3e-3
    ref2 = wan_vae.rms_norm(x.float(), gamma.float().view(C, 1, 1, 1))
    out2 = rmsnorm_silu_cl(_cl(x), gamma.to(DEV), silu=False).permute(3, 0, 1, 2).unsqueeze(0)
    assert rel_l2(out2, ref2) <= 3e-3
    up = upsample2x_cl(_cl(x)).permute(3, 0, 1, 2)                                   # [C,T,2H,2W]
    exp = F.interpolate(x[0].permute(1, 0, 2, 3).float(), scale_factor=(2.0, 2.0), mode="nearest-exact").permute(1, 0, 2, 3)
    assert torch.equal(up.float().cpu(), exp)


def test_softmax_rows_and_mid_attention():
    from apex_studio_b200.vae.wan import AutoencoderKLWan, WanVAEConfig, softmax_rows

    s = torch.randn(300, 520, device=DEV) * 4
    p = softmax_rows(s, 0.37)
    assert rel_l2(p, torch.softmax(s * 0.37, dim=-1)) <= 3e-3
    # whole attention block vs oracle
    C, T, H, W = 128, 2, 8, 12
    w = wan_vae.make_weights(base_dim=32, seed=7)
    vae = AutoencoderKLWan(WanVAEConfig(base_dim=32))
    vae.load_state_dict(w, device=DEV)
    x = torch.randn(1, C, T, H, W, generator=torch.Generator().manual_seed(2)).bfloat16()
    ref = wan_vae.attn_block(x.float(), w, "decoder.mid_block.attentions.0")
    out = vae._attn_block(_cl(x), "decoder.mid_block.attentions.0").permute(3, 0, 1, 2).unsqueeze(0)
    assert rel_l2(out, ref) <= 5e-3


def test_blend_tile_matches_reference_order():
    """2x2 tiles through b200_blend_tile vs the oracle's in-place blend/crop/clamp on the same bf16 tiles."""
    from apex_studio_b200.vae.wan import blend_tile

    g = torch.Generator().manual_seed(3)
    T, th, tw, stride, blend = 2, 32, 32, 24, 8
    shapes = [[(32, 32), (32, 20)], [(18, 32), (18, 20)]]
    tiles = [[(torch.randn(1, 3, T, h, w, generator=g) * 0.8).bfloat16() for (h, w) in row] for row in shapes]
    # oracle (bf16 tensors, in place, reference order)
    rows = [[t.clone() for t in row] for row in tiles]
    out_rows = []
    for i, row in enumerate(rows):
        res = []
        for j, tile in enumerate(row):
            if i > 0:
                tile = wan_vae._blend_v(rows[i - 1][j], tile, blend)
            if j > 0:
                tile = wan_vae._blend_h(row[j - 1], tile, blend)
            res.append(tile[:, :, :, :stride, :stride])
        out_rows.append(torch.cat(res, dim=-1))
    OH, OW = 24 + 18, 24 + 20
    expect = torch.clamp(torch.cat(out_rows, dim=3)[:, :, :, :OH, :OW], -1.0, 1.0)[0]
    dev_tiles = [[t[0].to(DEV).contiguous() for t in row] for row in tiles]
    frame = torch.zeros(3, T, OH, OW, device=DEV, dtype=torch.bfloat16)
    for i in range(2):
        for j in range(2):
            blend_tile(dev_tiles[i][j], dev_tiles[i - 1][j] if i > 0 else None, dev_tiles[i][j - 1] if j > 0 else None,
                       frame, blend, stride, i * stride, j * stride)
    assert torch.equal(frame.cpu(), expect)


@pytest.mark.parametrize("base_dim,shape", [(32, (16, 3, 12, 20)), (32, (16, 1, 8, 8)), (96, (16, 2, 18, 16)), (32, (16, 5, 32, 32))])
def test_c_entry_tile_decode_is_bit_identical_to_the_per_kernel_path(base_dim, shape):
    """b200_wan_vae_decode (ONE C call: the whole decoder launch sequence into a ping-pong workspace) vs the same kernels
    issued one by one from Python: identical bits, for ragged tiles, a single frame (no temporal upsampling) and width 96."""
    from apex_studio_b200.vae import AutoencoderKLWan, WanVAEConfig

    vae = AutoencoderKLWan(WanVAEConfig(base_dim=base_dim))
    vae.load_state_dict(wan_vae.make_weights(base_dim=base_dim, seed=3), device=DEV)
    z = torch.randn(*shape, generator=torch.Generator().manual_seed(sum(shape))).to(DEV, torch.bfloat16)
    a = vae.decode_tile(z)
    b = vae.decode_tile_py(z)
    assert a.shape == b.shape and torch.equal(a, b)
    assert torch.equal(vae.decode_tile(z), a)              # workspace reuse


@pytest.mark.parametrize("name,tiling,sub", [("untiled", False, 2), ("tiled", True, 3)])
def test_vae_decode_vs_reference_golden(name, tiling, sub):
    from apex_studio_b200.vae import AutoencoderKLWan, WanVAEConfig

    gold = np.load(os.path.join(GOLDEN, "wan_vae.npz"))
    w = wan_vae.make_weights(base_dim=32, seed=7)
    vae = AutoencoderKLWan(WanVAEConfig(base_dim=32))
    vae.load_state_dict(w, device=DEV)
    lat = torch.from_numpy(gold[name + "_latents"])
    z = vae.denormalize_latents(lat).to(torch.bfloat16)         # base_engine.py:2040-2048
    if tiling:
        vae.enable_tiling()
    out = vae.decode(z.to(DEV), return_dict=False)[0]
    assert list(out.shape) == gold[name + "_shape"].tolist() and out.dtype == torch.bfloat16
    assert out.abs().max().item() <= 1.0
    got = out[..., ::sub, ::sub]
    exact = torch.from_numpy(gold[name + "_out_fp32"])           # reference, fp32
    ref16 = torch.from_numpy(gold[name + "_out_bf16"])           # reference, bf16 pipeline on CPU
    ref_err, our_err = rel_l2(ref16, exact), rel_l2(got, exact)
    assert our_err <= max(1e-2, 1.5 * ref_err), (our_err, ref_err)
    assert rel_l2(got, ref16) <= 3e-2


@pytest.mark.parametrize("T,H,W", [(1, 1, 1), (3, 17, 29), (5, 64, 96)])
def test_frames_to_uint8_bit_exact(T, H, W):
    """b200_frames_to_uint8 vs the restated VideoProcessor.postprocess_video arithmetic: integer output, must be ==.
    Inputs cover the clamp on both sides, exact .5 ties of the final rounding and every bf16 value in [-1.5, 1.5]."""
    from apex_studio_b200.vae.wan import frames_to_uint8

    g = torch.Generator().manual_seed(T + H + W)
    v = (torch.randn(3, T, H, W, generator=g) * 0.8).bfloat16()
    flat = v.view(-1)
    special = torch.tensor([-1.0, 1.0, -1.5, 1.5, 0.0, -0.0, 1.0 / 255, 0.00390625, 0.99609375], dtype=torch.bfloat16)
    n = min(flat.numel(), special.numel())
    flat[:n] = special[:n]
    got = frames_to_uint8(v.to(DEV))
    assert got.dtype == torch.uint8 and tuple(got.shape) == (T, H, W, 3)
    assert np.array_equal(got.cpu().numpy(), wan_vae.frames_to_uint8(v))


def test_frames_to_uint8_all_bf16_values():
    from apex_studio_b200.vae.wan import frames_to_uint8

    bits = torch.arange(0, 65536, dtype=torch.int32).to(torch.int16)
    allv = bits.view(torch.bfloat16)
    allv = allv[torch.isfinite(allv.float())]
    pad = (-allv.numel()) % 3
    allv = torch.cat([allv, allv[:pad]])
    v = allv.view(3, 1, 1, -1).contiguous()
    got = frames_to_uint8(v.to(DEV))
    assert np.array_equal(got.cpu().numpy(), wan_vae.frames_to_uint8(v))


def test_preview_renderer_on_a_side_stream_equals_synchronous_decode():
    """Per-step preview decode (reference BaseEngine._render_step, base_engine.py:2927-2943): the frames delivered
    by the side-stream renderer equal a synchronous decode + hand-off of the same latents, and the denoise stream keeps going."""
    from apex_studio_b200 import denoise
    from apex_studio_b200.vae import AutoencoderKLWan, WanVAEConfig
    from apex_studio_b200.vae.wan import frames_to_uint8

    gold = np.load(os.path.join(GOLDEN, "wan_vae.npz"))
    vae = AutoencoderKLWan(WanVAEConfig(base_dim=32))
    vae.load_state_dict(wan_vae.make_weights(base_dim=32, seed=7), device=DEV)
    lat = torch.from_numpy(gold["untiled_latents"]).to(DEV)
    fn = denoise.wan_preview_decode_fn(vae)
    got = []
    pr = denoise.PreviewRenderer(fn, got.append, interval=2)
    total, lats = 5, []
    for i in range(total):
        cur = lat * (1.0 - 0.1 * i)
        lats.append(cur)
        pr.maybe_render(i, total, cur)
        cur = None
        torch.randn(512, 512, device=DEV) @ torch.randn(512, 512, device=DEV)        # "next step" work on the main stream
    pr.finish()
    assert pr.rendered_steps == [0, 1, 3] and len(got) == 3
    for frames, step in zip(got, pr.rendered_steps):
        ref = fn(lats[step]).cpu().numpy()
        assert frames.dtype == np.uint8 and frames.shape == ref.shape and (frames == ref).all()


@pytest.mark.parametrize("T,H,W,cin,cout", [(5, 18, 16, 96, 96), (3, 20, 24, 192, 192), (4, 9, 21, 384, 192), (2, 16, 16, 96, 32)])
def test_conv3d_fused_rmsnorm_silu_epilogue(T, H, W, cin, cout):
    """b200_conv3d_cl_norm_silu (conv1 -> norm2 -> SiLU of WanResidualBlock in one kernel) vs the two-kernel path: the epilogue
    normalises the bf16-rounded conv output exactly as the norm kernel reads it back; only the summation order of the squares
    differs, so results agree to a bf16 ulp on all but a few elements.  Also against fp32 math of the whole chain."""
    from apex_studio_b200.vae.wan import conv3d_cl, rmsnorm_silu_cl

    g = torch.Generator().manual_seed(T * 10 + cout)
    x = torch.randn(1, cin, T, H, W, generator=g).bfloat16()
    w = (torch.randn(cout, cin, 3, 3, 3, generator=g) * (27 * cin) ** -0.5).bfloat16()
    b = (torch.randn(cout, generator=g) * 0.1).bfloat16()
    gamma = (1 + 0.1 * torch.randn(cout, generator=g)).bfloat16().to(DEV)
    two = rmsnorm_silu_cl(conv3d_cl(_cl(x), _tap_major(w.float()), b.to(DEV), (3, 3, 3), cout), gamma)
    one = conv3d_cl(_cl(x), _tap_major(w.float()), b.to(DEV), (3, 3, 3), cout, norm_gamma=gamma)
    diff = (one.float() - two.float()).abs()
    ulp = two.float().abs().clamp_min(1e-3) * 2.0 ** -7
    assert (diff > ulp).float().mean().item() <= 2e-3 and rel_l2(one, two) <= 2e-3
    conv = wan_vae.causal_conv3d(x.float(), {"c.weight": w.float(), "c.bias": b.float()}, "c")[0]     # [C,T,H,W] fp32
    ref = torch.nn.functional.normalize(conv, dim=0) * cout ** 0.5 * gamma.float().cpu().view(-1, 1, 1, 1)
    ref = torch.nn.functional.silu(ref).permute(1, 2, 3, 0)
    assert rel_l2(one, ref) <= 6e-3


def test_conv3d_fused_norm_rejects_split_channel_tiles():
    """384 output channels run as two N tiles of 192: the pixel's channel vector is not in one accumulator row."""
    from apex_studio_b200 import _lib
    from apex_studio_b200.vae.wan import conv3d_cl, conv_norm_fusable

    assert conv_norm_fusable(96) and conv_norm_fusable(192) and not conv_norm_fusable(384)
    x = torch.randn(2, 8, 8, 64, device=DEV).bfloat16()
    w = torch.randn(27 * 384, 64, device=DEV).bfloat16()
    with pytest.raises((ValueError, RuntimeError)):
        conv3d_cl(x, w, None, (3, 3, 3), 384, norm_gamma=torch.ones(384, device=DEV).bfloat16())
