"""CPU ORACLE (test infrastructure -- never imported by the product package) for the HunyuanVideo-1.5 3D-VAE decode
(BASELINE.json configs[4] "... with tiled 3D-VAE decode").

Functional restatement, in plain torch, of the decode half of ``AutoencoderKLHunyuanVideo15`` (paths relative to
/root/reference/apps/api/src/vae/hunyuanvideo15/model.py):

    :52-90     HunyuanVideo15CausalConv3d (replicate padding, causal in time)  -> :func:`causal_conv`
    :93-127    HunyuanVideo15RMS_norm                                           -> :func:`rms_norm`
    :130-214   HunyuanVideo15AttnBlock (frame-causal single-head attention)     -> :func:`attn_block`
    :217-274   HunyuanVideo15Upsample (DCAE channel-to-space + shortcut)        -> :func:`upsample`
    :338-380   HunyuanVideo15ResnetBlock                                        -> :func:`resnet`
    :637-732   HunyuanVideo15Decoder3D                                          -> :func:`decoder`
    :974-1002, :1060-1119  blend_v / blend_h / tiled_decode                     -> :func:`tiled_decode`
    :1145-1150 denormalize_latents                                              -> :func:`denormalize_latents`

fp32 = exact math; bf16 = the reference's rounding points (every torch op rounds where the reference's does).
PINNING: oracle/make_golden.py golden_hy15vae runs the reference's own class -> tests/golden/hy15_vae.npz;
tests/test_oracle_hy15_vae.py compares (bit-exact, fp32 and bf16)."""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Weights = Dict[str, torch.Tensor]


def causal_conv(x: torch.Tensor, w: Weights, p: str) -> torch.Tensor:
    wt = w[p + ".conv.weight"]
    kt, kh, kw = wt.shape[2:]
    x = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2, kt - 1, 0), mode="replicate")
    return F.conv3d(x, wt, w[p + ".conv.bias"])


def rms_norm(x: torch.Tensor, gamma: torch.Tensor) -> torch.Tensor:
    return F.normalize(x, dim=1) * (x.shape[1] ** 0.5) * gamma + 0.0


def resnet(x: torch.Tensor, w: Weights, p: str) -> torch.Tensor:
    h = causal_conv(F.silu(rms_norm(x, w[p + ".norm1.gamma"])), w, p + ".conv1")
    h = causal_conv(F.silu(rms_norm(h, w[p + ".norm2.gamma"])), w, p + ".conv2")
    if (p + ".conv_shortcut.weight") in w:
        x = F.conv3d(x, w[p + ".conv_shortcut.weight"], w[p + ".conv_shortcut.bias"])
    return h + x


def causal_mask(n_frame: int, n_hw: int, dtype) -> torch.Tensor:
    seq = n_frame * n_hw
    frame = torch.arange(seq) // n_hw
    mask = torch.full((seq, seq), float("-inf"), dtype=dtype)
    mask[torch.arange(seq)[None, :] < ((frame + 1) * n_hw)[:, None]] = 0
    return mask


def attn_block(x: torch.Tensor, w: Weights, p: str) -> torch.Tensor:
    b, c, f, hh, ww = x.shape
    n = rms_norm(x, w[p + ".norm.gamma"])
    q, k, v = (F.conv3d(n, w[f"{p}.{nm}.weight"], w[f"{p}.{nm}.bias"]).reshape(b, c, f * hh * ww).permute(0, 2, 1).unsqueeze(1)
               for nm in ("to_q", "to_k", "to_v"))
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=causal_mask(f, hh * ww, q.dtype)[None].expand(b, -1, -1))
    o = o.squeeze(1).reshape(b, f, hh, ww, c).permute(0, 4, 1, 2, 3)
    return F.conv3d(o, w[p + ".proj_out.weight"], w[p + ".proj_out.bias"]) + x


def _rearrange(t: torch.Tensor, r1: int, r2: int = 2, r3: int = 2) -> torch.Tensor:
    b, pc, f, h, w_ = t.shape
    c = pc // (r1 * r2 * r3)
    return t.view(b, r1, r2, r3, c, f, h, w_).permute(0, 4, 5, 1, 6, 2, 7, 3).reshape(b, c, f * r1, h * r2, w_ * r3)


def upsample(x: torch.Tensor, w: Weights, p: str, temporal: bool) -> torch.Tensor:
    h = causal_conv(x, w, p + ".conv")
    factor = 8 if temporal else 4
    out_ch = h.shape[1] // factor
    repeats = factor * out_ch // x.shape[1]
    if temporal:
        h_first = _rearrange(h[:, :, :1], 1)
        h_first = h_first[:, : h_first.shape[1] // 2]
        h = torch.cat([h_first, _rearrange(h[:, :, 1:], 2)], dim=2)
        x_first = _rearrange(x[:, :, :1], 1).repeat_interleave(repeats // 2, dim=1)
        x_next = _rearrange(x[:, :, 1:], 2).repeat_interleave(repeats, dim=1)
        short = torch.cat([x_first, x_next], dim=2)
    else:
        h = _rearrange(h, 1)
        short = _rearrange(x.repeat_interleave(repeats, dim=1), 1)
    return h + short


def decoder(z: torch.Tensor, w: Weights, block_out_channels: Sequence[int], layers_per_block: int = 2,
            spatial_ratio: int = 16, temporal_ratio: int = 4) -> torch.Tensor:
    """:708-732 with ``block_out_channels`` already reversed (decoder order, e.g. (1024, 1024, 512, 256, 128))."""
    p = "decoder."
    x = causal_conv(z, w, p + "conv_in") + z.repeat_interleave(block_out_channels[0] // z.shape[1], dim=1)
    x = resnet(x, w, p + "mid_block.resnets.0")
    x = attn_block(x, w, p + "mid_block.attentions.0")
    x = resnet(x, w, p + "mid_block.resnets.1")
    for i in range(len(block_out_channels)):
        for j in range(layers_per_block + 1):
            x = resnet(x, w, f"{p}up_blocks.{i}.resnets.{j}")
        if i < np.log2(spatial_ratio) or i < np.log2(temporal_ratio):
            x = upsample(x, w, f"{p}up_blocks.{i}.upsamplers.0", temporal=bool(i < np.log2(temporal_ratio)))
    x = F.silu(rms_norm(x, w[p + "norm_out.gamma"]))
    return causal_conv(x, w, p + "conv_out")


def _blend(a: torch.Tensor, b: torch.Tensor, extent: int, dim: int) -> torch.Tensor:
    extent = min(a.shape[dim], b.shape[dim], extent)
    for y in range(extent):
        ia = [slice(None)] * 5
        ib = [slice(None)] * 5
        ia[dim], ib[dim] = -extent + y, y
        b[tuple(ib)] = a[tuple(ia)] * (1 - y / extent) + b[tuple(ib)] * (y / extent)
    return b


def tiled_decode(z: torch.Tensor, w: Weights, block_out_channels, tile_sample_min: int = 128, overlap: float = 0.25,
                 spatial_ratio: int = 16, **kw) -> torch.Tensor:
    """:1060-1119: latent tiles of tile_sample_min/16 at stride int(tile*(1-overlap)), row-major in-place blends, crop."""
    tl = tile_sample_min // spatial_ratio
    ov = int(tl * (1 - overlap))
    blend = int(tile_sample_min * overlap)
    limit = tile_sample_min - blend
    _, _, _, height, width = z.shape
    rows = []
    for i in range(0, height, ov):
        rows.append([decoder(z[:, :, :, i:i + tl, j:j + tl], w, block_out_channels, spatial_ratio=spatial_ratio, **kw)
                     for j in range(0, width, ov)])
    out_rows = []
    for i, row in enumerate(rows):
        res = []
        for j, tile in enumerate(row):
            if i > 0:
                tile = _blend(rows[i - 1][j], tile, blend, 3)
            if j > 0:
                tile = _blend(row[j - 1], tile, blend, 4)
            res.append(tile[:, :, :, :limit, :limit])
        out_rows.append(torch.cat(res, dim=-1))
    return torch.cat(out_rows, dim=-2)


def denormalize_latents(latents: torch.Tensor, scaling_factor: float = 1.03682, shift_factor=None) -> torch.Tensor:
    return latents / scaling_factor + shift_factor if shift_factor else latents / scaling_factor


def make_weights(block_out_channels: Sequence[int], latent_channels: int = 32, out_channels: int = 3, layers_per_block: int = 2,
                 spatial_ratio: int = 16, temporal_ratio: int = 4, seed: int = 7, dtype=torch.float32) -> Weights:
    """Synthetic decoder weights under the reference's state-dict keys (``block_out_channels`` in decoder order)."""
    g = torch.Generator().manual_seed(seed)
    w: Weights = {}

    def conv(name, cout, cin, k):
        fan = cin * k ** 3
        w[name + ".weight"] = torch.randn(cout, cin, k, k, k, generator=g) * fan ** -0.5
        w[name + ".bias"] = torch.randn(cout, generator=g) * 0.02

    def gamma(name, c):
        w[name] = (1 + 0.1 * torch.randn(c, generator=g)).view(c, 1, 1, 1)

    def res(p, cin, cout):
        gamma(p + ".norm1.gamma", cin)
        conv(p + ".conv1.conv", cout, cin, 3)
        gamma(p + ".norm2.gamma", cout)
        conv(p + ".conv2.conv", cout, cout, 3)
        if cin != cout:
            conv(p + ".conv_shortcut", cout, cin, 1)

    b = list(block_out_channels)
    p = "decoder."
    conv(p + "conv_in.conv", b[0], latent_channels, 3)
    res(p + "mid_block.resnets.0", b[0], b[0])
    gamma(p + "mid_block.attentions.0.norm.gamma", b[0])
    for nm in ("to_q", "to_k", "to_v", "proj_out"):
        conv(f"{p}mid_block.attentions.0.{nm}", b[0], b[0], 1)
    res(p + "mid_block.resnets.1", b[0], b[0])
    cur = b[0]
    for i, co in enumerate(b):
        for j in range(layers_per_block + 1):
            res(f"{p}up_blocks.{i}.resnets.{j}", cur if j == 0 else co, co)
        cur = co
        sp, tp = i < np.log2(spatial_ratio), i < np.log2(temporal_ratio)
        if sp or tp:
            up_out = b[i + 1]
            conv(f"{p}up_blocks.{i}.upsamplers.0.conv.conv", up_out * (8 if tp else 4), co, 3)
            cur = up_out
    gamma(p + "norm_out.gamma", b[-1])
    conv(p + "conv_out.conv", out_channels, b[-1], 3)
    return {k: v.to(dtype) for k, v in w.items()}
