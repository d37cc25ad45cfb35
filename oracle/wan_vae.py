"""CPU ORACLE (test infrastructure) for the Wan 3D-VAE decode path.

Functional torch restatement of what ``BaseEngine.vae_decode`` computes for the Wan 2.1/2.2-A14B VAE
(non-residual decoder; paths relative to /root/reference/apps/api/src/):

    engine/base_engine.py:2030-2059     vae_decode: denormalise, force tiling, decode   -> :func:`vae_decode`
    vae/wan/model.py:1649-1659          denormalize_latents                             -> :func:`denormalize_latents`
    vae/wan/model.py:1516-1623          tiled_decode (32x32 latent tiles, stride 24, blends, crop, clamp)
                                                                                        -> :func:`tiled_decode`
    vae/wan/model.py:1404-1422          blend_v / blend_h (in place, so order matters)  -> :func:`_blend_v/_blend_h`
    vae/wan/model.py:972-1021           WanDecoder3d.forward                            -> :func:`decoder_forward`
    vae/wan/model.py:389-441            WanResidualBlock                                -> :func:`res_block`
    vae/wan/model.py:461-490            WanAttentionBlock (single head, per frame)      -> :func:`attn_block`
    vae/wan/model.py:291-353            WanResample upsample2d / upsample3d             -> :func:`upsample`
    vae/wan/model.py:178-185            WanCausalConv3d                                 -> :func:`causal_conv3d`
    vae/wan/model.py:216-222            WanRMS_norm                                     -> :func:`rms_norm`

FORMULATION.  The reference decodes one latent frame at a time and carries the last CACHE_T = 2 frames of every
causal conv's input in ``feat_cache`` (:36).  Tracing that cache logic shows that each causal conv sees exactly
[0, 0, x_0, x_1, ...] -- i.e. the streaming decode equals ONE causal convolution over the whole time axis with two
zero frames of left padding -- and that ``upsample3d`` lets the first frame bypass ``time_conv`` ("Rep" sentinel,
:296-330) while frames 1.. go through a causal ``time_conv`` with zero history and are interleaved 2x.  The oracle is
written in that whole-sequence form; tests/test_oracle_vae.py checks it against the reference's own streaming
``AutoencoderKLWan.decode`` (golden vectors from oracle/make_golden.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Weights = Dict[str, torch.Tensor]

LATENTS_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632,
                -0.1922, -0.9497, 0.2503, -0.2921]
LATENTS_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382,
               1.1253, 2.8251, 1.916]


def causal_conv3d(x: torch.Tensor, w: Weights, prefix: str) -> torch.Tensor:
    """Conv3d with symmetric H/W padding (k-1)/2 and 2*pad_t = (kt-1) zero frames on the LEFT of time only."""
    weight, bias = w[prefix + ".weight"], w.get(prefix + ".bias")
    kt, kh, kw = weight.shape[2:]
    x = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2, kt - 1, 0))
    return F.conv3d(x, weight, bias)


def rms_norm(x: torch.Tensor, gamma: torch.Tensor) -> torch.Tensor:
    """F.normalize(x, dim=1) * sqrt(C) * gamma  (channel-first, no bias)."""
    return F.normalize(x, dim=1) * (x.shape[1] ** 0.5) * gamma


def res_block(x: torch.Tensor, w: Weights, p: str) -> torch.Tensor:
    h = causal_conv3d(x, w, p + ".conv_shortcut") if (p + ".conv_shortcut.weight") in w else x
    x = F.silu(rms_norm(x, w[p + ".norm1.gamma"]))
    x = causal_conv3d(x, w, p + ".conv1")
    x = F.silu(rms_norm(x, w[p + ".norm2.gamma"]))
    x = causal_conv3d(x, w, p + ".conv2")
    return x + h


def attn_block(x: torch.Tensor, w: Weights, p: str) -> torch.Tensor:
    b, c, t, hh, ww = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, hh, ww)
    y = rms_norm(y, w[p + ".norm.gamma"])
    qkv = F.conv2d(y, w[p + ".to_qkv.weight"], w[p + ".to_qkv.bias"])
    qkv = qkv.reshape(b * t, 1, c * 3, -1).permute(0, 1, 3, 2).contiguous()
    q, k, v = qkv.chunk(3, dim=-1)
    y = F.scaled_dot_product_attention(q, k, v)
    y = y.squeeze(1).permute(0, 2, 1).reshape(b * t, c, hh, ww)
    y = F.conv2d(y, w[p + ".proj.weight"], w[p + ".proj.bias"])
    y = y.view(b, t, c, hh, ww).permute(0, 2, 1, 3, 4)
    return y + x


def upsample(x: torch.Tensor, w: Weights, p: str, temporal: bool) -> torch.Tensor:
    """upsample3d: frames 1.. -> causal time_conv (C -> 2C) -> interleave to 2x frames; then per frame
    nearest-exact 2x (done in fp32, cast back) + Conv2d(C -> C/2, 3, pad 1)."""
    b, c, t, hh, ww = x.shape
    if temporal and t > 1:
        rest = causal_conv3d(x[:, :, 1:], w, p + ".time_conv")        # [b, 2c, t-1, h, w]
        rest = rest.reshape(b, 2, c, t - 1, hh, ww)
        rest = torch.stack((rest[:, 0], rest[:, 1]), 3).reshape(b, c, 2 * (t - 1), hh, ww)
        x = torch.cat([x[:, :, :1], rest], dim=2)
        t = x.shape[2]
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, hh, ww)
    y = F.interpolate(y.float(), scale_factor=(2.0, 2.0), mode="nearest-exact").type_as(y)
    y = F.conv2d(y, w[p + ".resample.1.weight"], w[p + ".resample.1.bias"], padding=1)
    return y.view(b, t, y.size(1), y.size(2), y.size(3)).permute(0, 2, 1, 3, 4)


def decoder_forward(z: torch.Tensor, w: Weights, *, num_res_blocks: int = 2,
                    temporal_upsample: Sequence[bool] = (True, True, False)) -> torch.Tensor:
    """post_quant_conv + WanDecoder3d over ALL latent frames of one (tile of a) latent: [B,zc,T,h,w] ->
    [B,3,1+4(T-1),8h,8w] (before clamp)."""
    x = causal_conv3d(z, w, "post_quant_conv")
    x = causal_conv3d(x, w, "decoder.conv_in")
    x = res_block(x, w, "decoder.mid_block.resnets.0")
    x = attn_block(x, w, "decoder.mid_block.attentions.0")
    x = res_block(x, w, "decoder.mid_block.resnets.1")
    n_up = len(temporal_upsample) + 1
    for i in range(n_up):
        for j in range(num_res_blocks + 1):
            x = res_block(x, w, f"decoder.up_blocks.{i}.resnets.{j}")
        if i != n_up - 1:
            x = upsample(x, w, f"decoder.up_blocks.{i}.upsamplers.0", temporal_upsample[i])
    x = F.silu(rms_norm(x, w["decoder.norm_out.gamma"]))
    return causal_conv3d(x, w, "decoder.conv_out")


def _blend_v(a, b, extent):
    extent = min(a.shape[-2], b.shape[-2], extent)
    for y in range(extent):
        b[:, :, :, y, :] = a[:, :, :, -extent + y, :] * (1 - y / extent) + b[:, :, :, y, :] * (y / extent)
    return b


def _blend_h(a, b, extent):
    extent = min(a.shape[-1], b.shape[-1], extent)
    for x in range(extent):
        b[:, :, :, :, x] = a[:, :, :, :, -extent + x] * (1 - x / extent) + b[:, :, :, :, x] * (x / extent)
    return b


def tile_grid(height: int, width: int, tile_min: int = 32, stride: int = 24) -> List[Tuple[int, int, int, int]]:
    """Latent-space tiles in the reference's order: (i, j, h, w) for i in range(0,H,stride), j in range(0,W,stride)."""
    return [(i, j, min(tile_min, height - i), min(tile_min, width - j))
            for i in range(0, height, stride) for j in range(0, width, stride)]


def tiled_decode(z: torch.Tensor, w: Weights, *, tile_sample_min: int = 256, tile_sample_stride: int = 192,
                 spatial_ratio: int = 8, decode_tile=None, **dec_kw) -> torch.Tensor:
    """vae/wan/model.py:1516-1623 (patch_size None).  ``decode_tile`` lets a test substitute the CUDA tile decoder."""
    _, _, _, height, width = z.shape
    tmin, tstride = tile_sample_min // spatial_ratio, tile_sample_stride // spatial_ratio
    blend = tile_sample_min - tile_sample_stride
    dec = decode_tile or (lambda t: decoder_forward(t, w, **dec_kw))
    rows = []
    for i in range(0, height, tstride):
        rows.append([dec(z[:, :, :, i:i + tmin, j:j + tmin]) for j in range(0, width, tstride)])
    out_rows = []
    for i, row in enumerate(rows):
        res = []
        for j, tile in enumerate(row):
            if i > 0:
                tile = _blend_v(rows[i - 1][j], tile, blend)
            if j > 0:
                tile = _blend_h(row[j - 1], tile, blend)
            res.append(tile[:, :, :, :tile_sample_stride, :tile_sample_stride])
        out_rows.append(torch.cat(res, dim=-1))
    out = torch.cat(out_rows, dim=3)[:, :, :, :height * spatial_ratio, :width * spatial_ratio]
    return torch.clamp(out, min=-1.0, max=1.0)


def decode(z: torch.Tensor, w: Weights, *, use_tiling: bool = True, tile_sample_min: int = 256,
           tile_sample_stride: int = 192, **dec_kw) -> torch.Tensor:
    """AutoencoderKLWan._decode (:1333-1375): tiled when the latent exceeds one tile, else whole-frame."""
    tmin = tile_sample_min // 8
    if use_tiling and (z.shape[-1] > tmin or z.shape[-2] > tmin):
        return tiled_decode(z, w, tile_sample_min=tile_sample_min, tile_sample_stride=tile_sample_stride, **dec_kw)
    return torch.clamp(decoder_forward(z, w, **dec_kw), min=-1.0, max=1.0)


def denormalize_latents(latents: torch.Tensor, mean=LATENTS_MEAN, std=LATENTS_STD) -> torch.Tensor:
    m = torch.tensor(mean).view(1, -1, 1, 1, 1).to(latents.device, latents.dtype)
    s = 1.0 / torch.tensor(std).view(1, -1, 1, 1, 1).to(latents.device, latents.dtype)
    return latents / s + m


def vae_decode(latents: torch.Tensor, w: Weights, dtype=torch.float32, **kw) -> torch.Tensor:
    """base_engine.py:2030-2059: denormalise in the latents' dtype (fp32), cast to the VAE dtype, tiled decode."""
    return decode(denormalize_latents(latents).to(dtype), w, use_tiling=True, **kw)


def frames_to_uint8(video: torch.Tensor):
    """BaseEngine._tensor_to_frames (engine/base_engine.py:2945-2949) -> diffusers VideoProcessor.postprocess_video
    (un-vendored dependency, restated from the published code: ``denormalize`` = ``(x * 0.5 + 0.5).clamp(0, 1)`` in the
    tensor's dtype, ``pt_to_numpy`` = ``.cpu().permute(0, 2, 3, 1).float().numpy()``, ``numpy_to_pil`` =
    ``(x * 255).round().astype("uint8")``), up to the uint8 arrays PIL wraps.  PARITY UNPINNED for this function: the
    reference holds no test or golden vector for it.
    video [3, T, H, W] (the dtype the VAE returned, bf16 on CUDA) -> numpy uint8 [T, H, W, 3]."""
    x = video.permute(1, 0, 2, 3)                       # postprocess_video: batch_vid = video[b].permute(1, 0, 2, 3)
    x = (x * 0.5 + 0.5).clamp(0, 1)                     # VaeImageProcessor.denormalize (dtype of the tensor)
    x = x.cpu().permute(0, 2, 3, 1).float().numpy()     # pt_to_numpy
    return (x * 255).round().astype("uint8")            # numpy_to_pil


def make_weights(*, base_dim: int = 96, z_dim: int = 16, dim_mult=(1, 2, 4, 4), num_res_blocks: int = 2,
                 temporal_upsample=(True, True, False), seed: int = 7, dtype=torch.float32) -> Weights:
    """Synthetic decoder weights with the reference's state-dict keys: conv weights ~ N(0, 1/fan_in) so activations
    keep O(1) scale through ~35 layers, biases N(0, 0.02^2), gammas 1 + N(0, 0.02^2)."""
    g = torch.Generator().manual_seed(seed)
    w: Weights = {}

    def conv(name, cout, cin, *k):
        fan = cin
        for kk in k:
            fan *= kk
        w[name + ".weight"] = torch.randn(cout, cin, *k, generator=g) * fan ** -0.5
        w[name + ".bias"] = torch.randn(cout, generator=g) * 0.02

    def gamma(name, c, nd):
        w[name] = (1 + torch.randn(c, generator=g) * 0.02).view(c, *([1] * nd))

    def res(p, cin, cout):
        gamma(p + ".norm1.gamma", cin, 3)
        conv(p + ".conv1", cout, cin, 3, 3, 3)
        gamma(p + ".norm2.gamma", cout, 3)
        conv(p + ".conv2", cout, cout, 3, 3, 3)
        if cin != cout:
            conv(p + ".conv_shortcut", cout, cin, 1, 1, 1)

    dims = [base_dim * u for u in [dim_mult[-1]] + list(dim_mult[::-1])]
    conv("post_quant_conv", z_dim, z_dim, 1, 1, 1)
    conv("decoder.conv_in", dims[0], z_dim, 3, 3, 3)
    res("decoder.mid_block.resnets.0", dims[0], dims[0])
    gamma("decoder.mid_block.attentions.0.norm.gamma", dims[0], 2)
    conv("decoder.mid_block.attentions.0.to_qkv", dims[0] * 3, dims[0], 1, 1)
    conv("decoder.mid_block.attentions.0.proj", dims[0], dims[0], 1, 1)
    res("decoder.mid_block.resnets.1", dims[0], dims[0])
    n_up = len(dim_mult)
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        if i > 0:
            cin = cin // 2
        cur = cin
        for j in range(num_res_blocks + 1):
            res(f"decoder.up_blocks.{i}.resnets.{j}", cur, cout)
            cur = cout
        if i != n_up - 1:
            p = f"decoder.up_blocks.{i}.upsamplers.0"
            conv(p + ".resample.1", cout // 2, cout, 3, 3)
            if temporal_upsample[i]:
                conv(p + ".time_conv", cout * 2, cout, 3, 1, 1)
    gamma("decoder.norm_out.gamma", dims[-1], 3)
    conv("decoder.conv_out", 3, dims[-1], 3, 3, 3)
    return {k: v.to(dtype) for k, v in w.items()}
