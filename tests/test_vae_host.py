"""Host-side logic of the VAE path that needs no GPU: the tap-major weight layout consumed by b200_conv3d_cl,
channel padding, tile grid, config surface."""
import pytest
import torch
import torch.nn.functional as F

import wan_vae
from apex_studio_b200.vae.wan import AutoencoderKLWan, WanVAEConfig


def _implicit_gemm_cpu(x_cl, w_tap, bias, taps, cout):
    """Pure-torch emulation of what the kernel computes from its operands: for every tap, a shifted window of the
    channels-last input (zero outside, KT-1 frames of causal left pad) times the tap's [Cout, Cin] matrix."""
    T, H, W, cin = x_cl.shape
    kt, kh, kw = taps
    xp = F.pad(x_cl, (0, 0, kw // 2, kw // 2, kh // 2, kh // 2, kt - 1, 0))
    y = torch.zeros(T, H, W, cout)
    w3 = w_tap.view(kt * kh * kw, cout, cin)
    for a in range(kt):
        for b in range(kh):
            for c in range(kw):
                win = xp[a:a + T, b:b + H, c:c + W]
                y += win @ w3[(a * kh + b) * kw + c].t()
    return y + bias


@pytest.mark.parametrize("taps", [(3, 3, 3), (1, 3, 3), (3, 1, 1)])
def test_tap_major_layout_equals_causal_conv(taps):
    g = torch.Generator().manual_seed(1)
    cin, cout, T, H, W = 8, 6, 4, 5, 7
    x = torch.randn(1, cin, T, H, W, generator=g)
    w = torch.randn(cout, cin, *taps, generator=g)
    b = torch.randn(cout, generator=g)
    ref = wan_vae.causal_conv3d(x, {"c.weight": w, "c.bias": b}, "c")[0].permute(1, 2, 3, 0)
    got = _implicit_gemm_cpu(x[0].permute(1, 2, 3, 0), AutoencoderKLWan._tap_major(w), b, taps, cout)
    assert torch.allclose(got, ref, atol=1e-4)


def test_channel_padding_and_config_surface():
    w = wan_vae.make_weights(base_dim=32, seed=7)
    vae = AutoencoderKLWan(WanVAEConfig(base_dim=32))
    vae.load_state_dict(w, device="cpu")
    pq = vae.w["post_quant_conv.weight"]
    assert pq.shape == (32, 64) and pq[16:].abs().max() == 0 and pq[:, 16:].abs().max() == 0
    assert vae.w["decoder.conv_in.weight"].shape == (27 * 128, 32)
    assert vae.w["decoder.conv_in.weight"].view(27, 128, 32)[:, :, 16:].abs().max() == 0
    assert vae.w["decoder.conv_out.weight"].shape == (27 * 16, 32)
    assert vae.w["decoder.conv_out.weight"].view(27, 16, 32)[:, 3:].abs().max() == 0
    assert vae.config.z_dim == 16 and vae.config.scale_factor_spatial == 8 and vae.config.scale_factor_temporal == 4
    assert vae.dtype == torch.bfloat16 and vae.dims == [128, 128, 128, 64, 32]
    assert vae.tile_grid(90, 160) == [(i, j) for (i, j, _, _) in wan_vae.tile_grid(90, 160)]
    lat = torch.randn(1, 16, 2, 4, 4)
    assert torch.equal(vae.denormalize_latents(lat), wan_vae.denormalize_latents(lat))
    with pytest.raises(ValueError):
        AutoencoderKLWan(WanVAEConfig(is_residual=True))
    with pytest.raises(RuntimeError, match="weights not loaded"):
        AutoencoderKLWan().decode(torch.zeros(1, 16, 1, 4, 4))


def test_frames_to_uint8_oracle_layout_and_rounding():
    """The oracle's restatement of VideoProcessor.postprocess_video: layout [T,H,W,3], bf16 x*0.5+0.5, clamp,
    round-half-even.  (Parity unpinned in the reference; pinned here to hand-computed values.)"""
    v = torch.tensor([-1.0, 1.0, 0.0, -2.0, 3.0, 1.0 / 255 * 2 - 1], dtype=torch.bfloat16).view(3, 1, 1, 2)
    out = wan_vae.frames_to_uint8(v)
    assert out.shape == (1, 1, 2, 3) and out.dtype.name == "uint8"
    # channel-major input: c0 = (-1, 1), c1 = (0, -2), c2 = (3, ~-0.992)
    assert out[0, 0, 0].tolist() == [0, 128, 255]            # 0.5 * 255 = 127.5 -> half-even -> 128
    assert out[0, 0, 1].tolist() == [255, 0, 1]


def test_frames_to_uint8_rejects_cpu_tensor():
    from apex_studio_b200.vae.wan import frames_to_uint8

    with pytest.raises(ValueError):
        frames_to_uint8(torch.zeros(3, 1, 2, 2, dtype=torch.bfloat16))


def test_fused_norm_rule_and_tile_launch_count(monkeypatch):
    """Which residual blocks run conv1 + norm2 + SiLU as one kernel (the rule shared by csrc/wan_vae.cu and vae/wan.py), and the
    launch count the C entry reports for a production tile (base_dim 96, 21 latent frames)."""
    from apex_studio_b200.vae.wan import conv_norm_fusable

    assert conv_norm_fusable(96) and conv_norm_fusable(192) and conv_norm_fusable(32)
    assert not conv_norm_fusable(384) and not conv_norm_fusable(100)     # two N tiles / not a multiple of 16
    vae = AutoencoderKLWan().init_random_weights("cpu", seed=7)
    res = [k for k in vae.w if k.endswith(".conv1.bias") and ".resnets." in k]
    assert len(res) == 14
    fused = sum(1 for k in res if vae.w[k].numel() <= 256)
    shortcuts = sum(1 for k in vae.w if k.endswith("conv_shortcut.weight"))
    temporal = sum(1 for t in vae.temporal_upsample if t)
    T = 21
    # post_quant + conv_in + conv_out + norm_out | residual blocks | shortcuts | mid attention (4 + 3 per frame) | 3 upsamplers
    expect = 4 + (14 * 4 - fused) + shortcuts + (4 + 3 * T) + 3 * 2 + temporal
    assert vae._tile_launches(T) == expect == 130
    monkeypatch.setenv("B200_VAE_FUSE_NORM", "0")
    assert not conv_norm_fusable(96)
    assert vae._tile_launches(T) == expect + fused
