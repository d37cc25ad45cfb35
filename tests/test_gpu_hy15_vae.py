"""GPU parity tests of the HunyuanVideo-1.5 VAE decode through the C ABI, against the CPU oracle
(oracle/hy15_vae.py, pinned to the reference's own AutoencoderKLHunyuanVideo15) and the reference's golden vectors.
Bars: data-movement kernels (replicate-pad gather without norm, DCAE rearrangement + shortcut) bit-exact; norm / softmax
kernels rel-L2 <= 4e-3 vs fp32 math; decode rel-L2 vs the exact fp32 reference output <= max(1e-3, 1.5 x the reference's
own bf16 error) and <= 2e-2 vs the reference's bf16 output."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import hy15_vae
from test_gpu_parity import rel_l2
from test_oracle_hy15_vae import CASES, CH, load

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cl(x):      # [1,C,T,H,W] -> channels-last [T,H,W,C]
    return x[0].permute(1, 2, 3, 0).contiguous()


def cf(x):      # [T,H,W,C] -> [1,C,T,H,W]
    return x.permute(3, 0, 1, 2).unsqueeze(0)


def test_pad_norm_silu_vs_torch():
    from apex_studio_b200.vae.hunyuanvideo15 import pad_norm_silu_cl

    torch.manual_seed(0)
    for C in (32, 128, 1024):
        x = (torch.randn(1, C, 3, 5, 6) * 2).bfloat16()
        g = (1 + 0.1 * torch.randn(C)).bfloat16()
        xp = F.pad(x.float(), (1, 1, 1, 1, 2, 0), mode="replicate")
        out = pad_norm_silu_cl(cl(x).to(DEV), None, False)
        assert torch.equal(cf(out.cpu()).float(), xp)                                    # pure gather: bit-exact
        ref = F.silu(F.normalize(xp, dim=1) * C ** 0.5 * g.float().view(1, C, 1, 1, 1))
        out = pad_norm_silu_cl(cl(x).to(DEV), g.to(DEV), True)
        assert tuple(out.shape) == (5, 7, 8, C) and rel_l2(cf(out), ref) <= 4e-3
        ref = F.normalize(x.float(), dim=1) * C ** 0.5 * g.float().view(1, C, 1, 1, 1)
        out = pad_norm_silu_cl(cl(x).to(DEV), g.to(DEV), False, pads=(0, 0, 0))
        assert rel_l2(cf(out), ref) <= 4e-3


@pytest.mark.parametrize("temporal,cin,cout", [(True, 64, 64), (True, 128, 64), (False, 64, 32), (False, 128, 128)])
def test_dcae_upsample_bit_exact_vs_reference_rearrangement(temporal, cin, cout):
    """h + shortcut of HunyuanVideo15Upsample.forward (model.py:249-274), the conv output h given."""
    from apex_studio_b200.vae.hunyuanvideo15 import dcae_upsample_cl

    torch.manual_seed(1)
    factor = 8 if temporal else 4
    x = torch.randn(1, cin, 3, 4, 5).bfloat16()
    h = torch.randn(1, cout * factor, 3, 4, 5).bfloat16()
    repeats = factor * cout // cin
    if temporal:
        hf = hy15_vae._rearrange(h[:, :, :1], 1)
        hh = torch.cat([hf[:, : hf.shape[1] // 2], hy15_vae._rearrange(h[:, :, 1:], 2)], dim=2)
        short = torch.cat([hy15_vae._rearrange(x[:, :, :1], 1).repeat_interleave(repeats // 2, dim=1),
                           hy15_vae._rearrange(x[:, :, 1:], 2).repeat_interleave(repeats, dim=1)], dim=2)
    else:
        hh = hy15_vae._rearrange(h, 1)
        short = hy15_vae._rearrange(x.repeat_interleave(repeats, dim=1), 1)
    ref = hh + short
    out = dcae_upsample_cl(cl(h).to(DEV), cl(x).to(DEV), cout, temporal)
    assert torch.equal(cf(out.cpu()), ref)


def test_block_causal_softmax_vs_torch():
    from apex_studio_b200.vae.hunyuanvideo15 import softmax_rows_block_causal

    torch.manual_seed(2)
    n_frame, n_hw = 5, 24
    s = torch.randn(n_frame * n_hw, n_frame * n_hw) * 4
    ref = torch.softmax(s * 0.25 + hy15_vae.causal_mask(n_frame, n_hw, torch.float32), dim=-1)
    out = softmax_rows_block_causal(s.to(DEV), 0.25, n_hw).float().cpu()
    assert (out[ref == 0] == 0).all() and rel_l2(out, ref) <= 4e-3


def _vae(channels=CH):
    from apex_studio_b200.vae import AutoencoderKLHunyuanVideo15, HunyuanVideo15VAEConfig

    return AutoencoderKLHunyuanVideo15(HunyuanVideo15VAEConfig(block_out_channels=tuple(reversed(channels))))


@pytest.mark.parametrize("name", list(CASES))
def test_vae_decode_vs_reference_golden(name):
    g = load()
    tiling, sub = CASES[name]
    vae = _vae()
    vae.load_state_dict(hy15_vae.make_weights(CH, seed=7), device=DEV)
    if tiling:
        vae.enable_tiling()
    z = torch.from_numpy(g[name + "_latents"]).to(DEV, torch.bfloat16)
    y = vae.decode(z, return_dict=False)[0]
    assert tuple(y.shape) == tuple(g[name + "_shape"]) and y.dtype == torch.bfloat16
    got = y[..., ::sub, ::sub]
    exact, ref16 = torch.from_numpy(g[f"{name}_out_fp32"]), torch.from_numpy(g[f"{name}_out_bf16"])
    ours, theirs = rel_l2(got, exact), rel_l2(ref16, exact)
    assert ours <= max(1e-3, 1.5 * theirs), (ours, theirs)
    assert rel_l2(got, ref16) <= 2e-2


def test_vae_production_width_tile_vs_exact_oracle():
    """One 8x8-latent tile, 2 latent frames, at the production widths (1024, 1024, 512, 256, 128): mid attention with head
    dim 1024, the 8192-channel upsample conv, the 128-channel output stage."""
    ch = (1024, 1024, 512, 256, 128)
    w32 = hy15_vae.make_weights(ch, seed=3)
    vae = _vae(ch)
    vae.load_state_dict(w32, device=DEV)
    z = torch.randn(1, 32, 2, 8, 8, generator=torch.Generator().manual_seed(5))
    y = vae.decode(z.to(DEV, torch.bfloat16), return_dict=False)[0]
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    exact = hy15_vae.decoder(z.bfloat16().float(), w32, ch)
    bf = hy15_vae.decoder(z.bfloat16(), {k: v.bfloat16() for k, v in w32.items()}, ch)
    assert tuple(y.shape) == (1, 3, 5, 128, 128) and torch.isfinite(y).all()
    ours, theirs = rel_l2(y, exact), rel_l2(bf, exact)
    assert ours <= max(1e-3, 1.5 * theirs), (ours, theirs)


@pytest.mark.parametrize("T,H,W,cin,cout", [(3, 8, 8, 32, 64), (2, 5, 7, 128, 48), (4, 16, 12, 64, 256)])
def test_conv3d_on_prepadded_input_vs_torch_replicate_conv(T, H, W, cin, cout):
    """HunyuanVideo15CausalConv3d (model.py:52-90) = replicate pad (2 frames in front, 1 pixel around) + valid 3x3x3 conv:
    gather kernel + tcgen05 implicit GEMM on the pre-padded tensor; rel-L2 <= 4e-3 vs fp32 math of the same bf16 operands."""
    from apex_studio_b200.vae.hunyuanvideo15 import conv3d_cl_padded, pad_norm_silu_cl
    from apex_studio_b200.vae.wan import AutoencoderKLWan

    torch.manual_seed(T * 100 + cin)
    x = torch.randn(1, cin, T, H, W).bfloat16()
    w = (torch.randn(cout, cin, 3, 3, 3) * (cin * 27) ** -0.5).bfloat16()
    b = (torch.randn(cout) * 0.1).bfloat16()
    res = torch.randn(1, cout, T, H, W).bfloat16()
    ref = F.conv3d(F.pad(x.float(), (1, 1, 1, 1, 2, 0), mode="replicate"), w.float(), b.float())
    xp = pad_norm_silu_cl(cl(x).to(DEV), None, False)
    out = conv3d_cl_padded(xp, AutoencoderKLWan._tap_major(w.float()).to(DEV, torch.bfloat16), b.to(DEV), cout)
    assert tuple(out.shape) == (T, H, W, cout) and rel_l2(cf(out), ref) <= 4e-3
    out = conv3d_cl_padded(xp, AutoencoderKLWan._tap_major(w.float()).to(DEV, torch.bfloat16), b.to(DEV), cout, residual=cl(res).to(DEV))
    assert rel_l2(cf(out), ref + res.float()) <= 4e-3


def test_blend_without_clamp_matches_reference_order():
    """AutoencoderKLHunyuanVideo15.tiled_decode (model.py:1060-1119): in-place blend_v then blend_h in bf16, crop, NO clamp."""
    from apex_studio_b200.vae.hunyuanvideo15 import blend_tile_noclamp

    torch.manual_seed(4)
    up, left, tile = (torch.randn(1, 3, 2, 16, 16).bfloat16() * 2 for _ in range(3))
    ref = hy15_vae._blend(left.clone(), hy15_vae._blend(up.clone(), tile.clone(), 4, 3), 4, 4)
    frame = torch.zeros(3, 2, 12, 12, dtype=torch.bfloat16, device=DEV)
    t = tile[0].to(DEV).contiguous()
    blend_tile_noclamp(t, up[0].to(DEV).contiguous(), left[0].to(DEV).contiguous(), frame, 4, 12, 0, 0)
    assert torch.equal(t.cpu(), ref[0]) and torch.equal(frame.cpu(), ref[0][..., :12, :12])
    assert ref.abs().max() > 1.0                              # values beyond [-1, 1] survive: no clamp
