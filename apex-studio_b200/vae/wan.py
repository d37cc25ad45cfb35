"""Wan 3-D causal VAE -- decode path on the B200 kernels.

Host-side mirror of the decode half of the reference's ``AutoencoderKLWan`` (apps/api/src/vae/wan/model.py:1083;
``decode`` :1378, ``_decode`` :1333-1375, ``tiled_decode`` :1516-1623, ``denormalize_latents`` :1649-1659) for the
non-residual decoder the Wan 2.1 / 2.2-A14B pipelines use (``WanDecoder3d`` :881-1021): same state-dict keys, the
``decode(z, return_dict=False)[0]`` / ``denormalize_latents`` / ``enable_tiling`` / ``.dtype`` / ``.config`` surface
``BaseEngine.vae_decode`` (engine/base_engine.py:2030-2059) relies on, so it can be registered as
``VAE_REGISTRY["wan_b200"]``.

Design (DESIGN.md section "VAE"):
* activations are channels-last bf16 ``[T, H, W, C]`` so that every causal conv is an implicit GEMM whose A operand
  is a TMA box with hardware zero fill for the spatial halo and the causal left pad (csrc/conv.cu);
* the reference's frame-by-frame ``feat_cache`` streaming equals one causal convolution over the whole time axis
  (proved against the reference in tests/test_oracle_vae.py), so a tile is decoded for ALL frames at once:
  ~60 launches per tile instead of ~35 convs x 21 frames;
* the 32x32-latent / stride-24 tiling, the in-place blend order and the crop are reproduced exactly (tiles see zero
  padding at their borders in the reference, so decoding untiled would NOT give the same pixels);
* tiles are independent until the blend: with N GPUs they are dealt by cost (longest-processing-time first: the edge tiles are
  smaller) and reassembled by ONE all-gather.
"""
from __future__ import annotations

import os

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .. import _lib, ops
from ..parallel import ParallelContext, deal_lpt


@dataclass
class WanVAEConfig:
    """Constructor arguments of the reference class that matter for decode (vae/wan/model.py:1100-1160)."""
    base_dim: int = 96
    decoder_base_dim: Optional[int] = None
    z_dim: int = 16
    dim_mult: Sequence[int] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    temperal_downsample: Sequence[bool] = (False, True, True)
    latents_mean: Sequence[float] = (-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134,
                                     -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921)
    latents_std: Sequence[float] = (2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526,
                                    2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.916)
    is_residual: bool = False
    out_channels: int = 3
    patch_size: Optional[int] = None
    scale_factor_temporal: int = 4
    scale_factor_spatial: int = 8


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# ---------------------------------------------------------------------------------------------------------
# thin wrappers over the C ABI
# ---------------------------------------------------------------------------------------------------------
def conv_norm_fusable(cout: int) -> bool:
    """conv1 -> norm2 -> SiLU of a residual block as one kernel (b200_conv3d_cl_norm_silu): the conv must keep a pixel's
    channel vector in one N tile (cout <= 256); B200_VAE_FUSE_NORM=0 disables it (the rule of csrc/wan_vae.cu)."""
    return os.environ.get("B200_VAE_FUSE_NORM", "1") != "0" and cout <= 256 and cout % 16 == 0


def conv3d_cl(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], taps: Tuple[int, int, int], cout: int, *,
              residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, planar_channels: int = 0,
              interleave: bool = False, norm_gamma: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [T,H,W,Cin] channels-last bf16, w [taps*cout, Cin] tap-major.  Returns [T,H,W,cout] (or, with
    ``interleave``, writes frames 1.. of ``out`` [1+2T,H,W,cout/2]; or planar [planar_channels,T,H,W]).
    ``norm_gamma``: WanRMS_norm + SiLU of the conv output fused into the epilogue (only that is returned)."""
    T, H, W, cin = x.shape
    kt, kh, kw = taps
    if not x.is_contiguous():
        raise ValueError("conv3d_cl needs a contiguous channels-last input")
    if norm_gamma is not None:
        if residual is not None or planar_channels or interleave:
            raise ValueError("norm_gamma fuses the norm of a plain conv output (no residual / planar / interleave)")
        if out is None:
            out = torch.empty(T, H, W, cout, dtype=torch.bfloat16, device=x.device)
        lib = _lib.load()
        rc = lib.b200_conv3d_cl_norm_silu(x.data_ptr(), w.data_ptr(), None if bias is None else bias.data_ptr(),
                                          norm_gamma.data_ptr(), out.data_ptr(), T, H, W, cin, cout, kt, kh, kw, _stream())
        _lib.check(rc, "b200_conv3d_cl_norm_silu")
        ops._count()
        return out
    if planar_channels:
        if out is None:
            out = torch.empty(planar_channels, T, H, W, dtype=torch.bfloat16, device=x.device)
        mode, tmul, toff, csplit, cvalid = 1, 1, 0, cout, planar_channels
    elif interleave:
        if out is None:
            raise ValueError("interleave writes into an existing [1+2T,H,W,C/2] buffer")
        mode, tmul, toff, csplit, cvalid = 0, 2, 1, cout // 2, cout
    else:
        if out is None:
            out = torch.empty(T, H, W, cout, dtype=torch.bfloat16, device=x.device)
        mode, tmul, toff, csplit, cvalid = 0, 1, 0, cout, cout
    lib = _lib.load()
    rc = lib.b200_conv3d_cl(x.data_ptr(), w.data_ptr(), None if bias is None else bias.data_ptr(),
                            None if residual is None else residual.data_ptr(), out.data_ptr(), T, H, W, cin, cout, kt, kh,
                            kw, mode, tmul, toff, csplit, cvalid, _stream())
    _lib.check(rc, "b200_conv3d_cl")
    ops._count()
    return out


def rmsnorm_silu_cl(x: torch.Tensor, gamma: torch.Tensor, silu: bool = True, out: Optional[torch.Tensor] = None):
    C = x.shape[-1]
    if out is None:
        out = torch.empty_like(x)
    lib = _lib.load()
    rc = lib.b200_rmsnorm_silu_cl(x.data_ptr(), out.data_ptr(), gamma.data_ptr(), x.numel() // C, C, int(silu), _stream())
    _lib.check(rc, "b200_rmsnorm_silu_cl")
    ops._count()
    return out


def upsample2x_cl(x: torch.Tensor) -> torch.Tensor:
    T, H, W, C = x.shape
    out = torch.empty(T, 2 * H, 2 * W, C, dtype=x.dtype, device=x.device)
    lib = _lib.load()
    rc = lib.b200_upsample2x_cl(x.data_ptr(), out.data_ptr(), T, H, W, C, _stream())
    _lib.check(rc, "b200_upsample2x_cl")
    ops._count()
    return out


def softmax_rows(s: torch.Tensor, scale: float) -> torch.Tensor:
    rows, cols = s.shape
    p = torch.empty(rows, cols, dtype=torch.bfloat16, device=s.device)
    lib = _lib.load()
    rc = lib.b200_softmax_rows(s.data_ptr(), p.data_ptr(), rows, cols, s.stride(0), p.stride(0), float(scale), _stream())
    _lib.check(rc, "b200_softmax_rows")
    ops._count()
    return p


def blend_tile(tile: torch.Tensor, up: Optional[torch.Tensor], left: Optional[torch.Tensor], frame: torch.Tensor,
               blend: int, crop: int, y0: int, x0: int) -> None:
    """tile/up/left: planar [3, T, th, tw] bf16 (contiguous); frame: [3, T, OH, OW]."""
    planes = tile.shape[0] * tile.shape[1]
    th, tw = tile.shape[-2:]
    uh, uw = (up.shape[-2], up.shape[-1]) if up is not None else (0, 0)
    lh, lw = (left.shape[-2], left.shape[-1]) if left is not None else (0, 0)
    lib = _lib.load()
    rc = lib.b200_blend_tile(tile.data_ptr(), None if up is None else up.data_ptr(),
                             None if left is None else left.data_ptr(), frame.data_ptr(), planes, th, tw, uh, uw, lh, lw,
                             blend, min(crop, th), min(crop, tw), y0, x0, frame.shape[-2], frame.shape[-1], _stream())
    _lib.check(rc, "b200_blend_tile")
    ops._count()


def frames_to_uint8(video: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Frame hand-off after decode: video [3, T, H, W] bf16 in [-1, 1] -> uint8 [T, H, W, 3] on the device, ready for a
    single D2H copy into the encoder.  Same arithmetic as ``BaseEngine._tensor_to_frames``
    (engine/base_engine.py:2945-2949 -> diffusers ``VideoProcessor.postprocess_video``: bf16 ``x*0.5+0.5``, clamp,
    float ``*255``, round-half-even) -- bit-exact, but without the reference's permute / float / PIL passes on the host."""
    ops._require_cuda_bf16("video", video)
    if video.dim() != 4 or video.shape[0] != 3 or not video.is_contiguous():
        raise ValueError(f"video must be contiguous planar [3, T, H, W], got {tuple(video.shape)}")
    _, T, H, W = video.shape
    if out is None:
        out = torch.empty(T, H, W, 3, dtype=torch.uint8, device=video.device)
    elif out.dtype != torch.uint8 or tuple(out.shape) != (T, H, W, 3) or not out.is_contiguous():
        raise ValueError("out must be a contiguous uint8 [T, H, W, 3] tensor")
    lib = _lib.load()
    rc = lib.b200_frames_to_uint8(video.data_ptr(), out.data_ptr(), T, H, W, _stream())
    _lib.check(rc, "b200_frames_to_uint8")
    ops._count()
    return out


# ---------------------------------------------------------------------------------------------------------
# the decoder
# ---------------------------------------------------------------------------------------------------------
class AutoencoderKLWan:
    """Decode-only B200 implementation (see module docstring)."""

    def __init__(self, config: Optional[WanVAEConfig] = None, **kwargs):
        self.config = config or WanVAEConfig(**kwargs)
        if self.config.is_residual or self.config.patch_size is not None:
            raise ValueError("only the non-residual Wan 2.1 / 2.2-A14B VAE decoder is implemented (is_residual=False)")
        self.dtype = torch.bfloat16
        self.device = None
        self.w: Dict[str, torch.Tensor] = {}
        self.use_tiling = False
        self.tile_sample_min_height = self.tile_sample_min_width = 256
        self.tile_sample_stride_height = self.tile_sample_stride_width = 192
        self.spatial_compression_ratio = self.config.scale_factor_spatial
        c = self.config
        d = c.decoder_base_dim or c.base_dim
        self.dims = [d * u for u in [c.dim_mult[-1]] + list(c.dim_mult[::-1])]
        self.temporal_upsample = list(c.temperal_downsample[::-1])

    # -------------------------------------------------------------------------------- weights
    @staticmethod
    def _tap_major(w: torch.Tensor, cin_pad: int = 0, cout_pad: int = 0) -> torch.Tensor:
        """[Cout,Cin,kt,kh,kw] or [Cout,Cin,kh,kw] -> [taps*Cout', Cin'] (tap-major, zero padded channels)."""
        if w.dim() == 4:
            w = w.unsqueeze(2)
        cout, cin = w.shape[:2]
        if cin_pad > cin:
            w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, cin_pad - cin))
        if cout_pad > cout:
            w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, 0, 0, cout_pad - cout))
        taps = w.shape[2] * w.shape[3] * w.shape[4]
        return w.permute(2, 3, 4, 0, 1).reshape(taps * w.shape[0], w.shape[1]).contiguous()

    def load_state_dict(self, state: Dict[str, torch.Tensor], device="cuda", strict: bool = False):
        """Reference state dict (decoder.* and post_quant_conv.*; encoder keys are ignored) -> bf16 device tensors in
        the kernels' layouts.  Channel padding: conv_in consumes 32 channels (z_dim 16 + 16 zero channels produced
        by a zero-padded post_quant_conv), conv_out produces 16 (3 real)."""
        dev = torch.device(device)
        self.device = dev
        bf = torch.bfloat16
        w: Dict[str, torch.Tensor] = {}
        z = self.config.z_dim
        zp = max(32, (z + 31) // 32 * 32)
        self.z_pad = zp
        for k, v in state.items():
            if not (k.startswith("decoder.") or k.startswith("post_quant_conv.")):
                continue
            v = v.detach().to(dev, torch.float32)
            if k == "post_quant_conv.weight":
                # [zp, 64]: output channels padded to zp (zeros), input channels padded to one 64-wide K tile
                w[k] = torch.nn.functional.pad(v.reshape(z, z), (0, 64 - z, 0, zp - z)).to(bf).contiguous()
            elif k == "post_quant_conv.bias":
                w[k] = torch.nn.functional.pad(v, (0, zp - z)).to(bf).contiguous()
            elif k == "decoder.conv_in.weight":
                w[k] = self._tap_major(v, cin_pad=zp).to(bf)
            elif k == "decoder.conv_out.weight":
                w[k] = self._tap_major(v, cout_pad=16).to(bf)
            elif k == "decoder.conv_out.bias":
                w[k] = torch.nn.functional.pad(v, (0, 16 - v.numel())).to(bf).contiguous()
            elif k.endswith("conv_shortcut.weight") or k.endswith("to_qkv.weight") or k.endswith("proj.weight"):
                w[k] = v.reshape(v.shape[0], v.shape[1]).to(bf).contiguous()                              # 1x1 -> linear
            elif k.endswith(".weight") and v.dim() >= 4:
                w[k] = self._tap_major(v).to(bf)
            elif k.endswith("gamma"):
                w[k] = v.reshape(-1).to(bf).contiguous()
            else:
                w[k] = v.to(bf).contiguous()
        self.w = w
        self._cw = None
        return [], []

    def init_random_weights(self, device="cuda", seed: int = 7):
        """Synthetic decoder weights in the reference's state-dict format, generated on the device (bench.py; no
        checkpoints offline): conv weights N(0, 1/fan_in), biases N(0, 0.02^2), gammas 1 + N(0, 0.02^2)."""
        dev = torch.device(device)
        g = torch.Generator(device=dev).manual_seed(seed)
        c = self.config
        sd: Dict[str, torch.Tensor] = {}

        def conv(name, cout, cin, *k):
            fan = cin
            for kk in k:
                fan *= kk
            sd[name + ".weight"] = torch.randn(cout, cin, *k, generator=g, device=dev) * fan ** -0.5
            sd[name + ".bias"] = torch.randn(cout, generator=g, device=dev) * 0.02

        def gamma(name, ch):
            sd[name] = 1 + torch.randn(ch, generator=g, device=dev) * 0.02

        def res(p, cin, cout):
            gamma(p + ".norm1.gamma", cin)
            conv(p + ".conv1", cout, cin, 3, 3, 3)
            gamma(p + ".norm2.gamma", cout)
            conv(p + ".conv2", cout, cout, 3, 3, 3)
            if cin != cout:
                conv(p + ".conv_shortcut", cout, cin, 1, 1, 1)

        dims = self.dims
        conv("post_quant_conv", c.z_dim, c.z_dim, 1, 1, 1)
        conv("decoder.conv_in", dims[0], c.z_dim, 3, 3, 3)
        res("decoder.mid_block.resnets.0", dims[0], dims[0])
        gamma("decoder.mid_block.attentions.0.norm.gamma", dims[0])
        conv("decoder.mid_block.attentions.0.to_qkv", dims[0] * 3, dims[0], 1, 1)
        conv("decoder.mid_block.attentions.0.proj", dims[0], dims[0], 1, 1)
        res("decoder.mid_block.resnets.1", dims[0], dims[0])
        n_up = len(c.dim_mult)
        for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
            cur = cin // 2 if i > 0 else cin
            for j in range(c.num_res_blocks + 1):
                res(f"decoder.up_blocks.{i}.resnets.{j}", cur, cout)
                cur = cout
            if i != n_up - 1:
                p = f"decoder.up_blocks.{i}.upsamplers.0"
                conv(p + ".resample.1", cout // 2, cout, 3, 3)
                if self.temporal_upsample[i]:
                    conv(p + ".time_conv", cout * 2, cout, 3, 1, 1)
        gamma("decoder.norm_out.gamma", dims[-1])
        conv("decoder.conv_out", c.out_channels, dims[-1], 3, 3, 3)
        self.load_state_dict(sd, device=dev)
        return self

    # -------------------------------------------------------------------------------- reference API surface
    def enable_tiling(self, tile_sample_min_height=None, tile_sample_min_width=None, tile_sample_stride_height=None,
                      tile_sample_stride_width=None) -> None:
        self.use_tiling = True
        self.tile_sample_min_height = tile_sample_min_height or self.tile_sample_min_height
        self.tile_sample_min_width = tile_sample_min_width or self.tile_sample_min_width
        self.tile_sample_stride_height = tile_sample_stride_height or self.tile_sample_stride_height
        self.tile_sample_stride_width = tile_sample_stride_width or self.tile_sample_stride_width

    def disable_tiling(self) -> None:
        self.use_tiling = False

    def denormalize_latents(self, latents: torch.Tensor) -> torch.Tensor:
        """vae/wan/model.py:1649-1659 (in the latents' own dtype, fp32 in the pipeline)."""
        c = self.config
        mean = torch.tensor(c.latents_mean).view(1, c.z_dim, 1, 1, 1).to(latents.device, latents.dtype)
        std = 1.0 / torch.tensor(c.latents_std).view(1, c.z_dim, 1, 1, 1).to(latents.device, latents.dtype)
        return latents / std + mean

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    # -------------------------------------------------------------------------------- blocks
    def _res_block(self, x: torch.Tensor, p: str) -> torch.Tensor:
        w = self.w
        cout = w[p + ".conv1.bias"].numel()
        if (p + ".conv_shortcut.weight") in w:
            h = ops.linear(x, w[p + ".conv_shortcut.weight"], w[p + ".conv_shortcut.bias"])
        else:
            h = x
        n = rmsnorm_silu_cl(x, w[p + ".norm1.gamma"])
        if conv_norm_fusable(cout):     # conv1 -> norm2 -> SiLU in one kernel (same rule as csrc/wan_vae.cu)
            n = conv3d_cl(n, w[p + ".conv1.weight"], w[p + ".conv1.bias"], (3, 3, 3), cout, norm_gamma=w[p + ".norm2.gamma"])
        else:
            y = conv3d_cl(n, w[p + ".conv1.weight"], w[p + ".conv1.bias"], (3, 3, 3), cout)
            del n
            n = rmsnorm_silu_cl(y, w[p + ".norm2.gamma"], out=y)
        return conv3d_cl(n, w[p + ".conv2.weight"], w[p + ".conv2.bias"], (3, 3, 3), cout, residual=h)

    def _attn_block(self, x: torch.Tensor, p: str) -> torch.Tensor:
        """Single-head attention over the H*W positions of each frame (head dim = C), x += proj(attn)."""
        w = self.w
        T, H, W, C = x.shape
        N = H * W
        xn = rmsnorm_silu_cl(x, w[p + ".norm.gamma"], silu=False)
        wq, bq = w[p + ".to_qkv.weight"], w[p + ".to_qkv.bias"]
        scale = C ** -0.5
        # the weight GEMMs do not mix frames: q | k, V^T and the output projection run once over all T frames; only the
        # scores and P V are per frame (66 launches instead of 126; same order as csrc/wan_vae.cu::attn_block)
        xa = xn.view(T * N, C)
        qk = ops.linear(xa, wq[:2 * C], bq[:2 * C])                                        # [T*N, 2C]
        vT = ops.linear(wq[2 * C:], xa, bq[2 * C:], row_bias=True)                          # [C, T*N] = V^T of every frame
        o = torch.empty(T * N, C, dtype=x.dtype, device=x.device)
        for t in range(T):
            rows = slice(t * N, (t + 1) * N)
            s = ops.linear(qk[rows, :C], qk[rows, C:], None, epilogue=ops.EPI_BIAS_F32)     # [N, N] fp32 scores
            pm = softmax_rows(s, scale)
            ops.linear(pm, vT[:, rows], out=o[rows])                                        # [N, C]
        ops.linear(o, w[p + ".proj.weight"], w[p + ".proj.bias"], epilogue=ops.EPI_GATE_RES, out=x.view(T * N, C), gate=None)
        return x

    def _upsample(self, x: torch.Tensor, p: str, temporal: bool) -> torch.Tensor:
        w = self.w
        T, H, W, C = x.shape
        if temporal and T > 1:
            y = torch.empty(1 + 2 * (T - 1), H, W, C, dtype=x.dtype, device=x.device)
            y[0].copy_(x[0])
            conv3d_cl(x[1:], w[p + ".time_conv.weight"], w[p + ".time_conv.bias"], (3, 1, 1), 2 * C, out=y, interleave=True)
            x = y
        u = upsample2x_cl(x)
        return conv3d_cl(u, w[p + ".resample.1.weight"], w[p + ".resample.1.bias"], (1, 3, 3), C // 2)

    # -------------------------------------------------------------------------------- one C call per tile
    def _c_weights(self) -> "_lib.WanVaeWeights":
        """The B200WanVaeWeights table (include/apex_b200.h) over the resident weights; built once per load."""
        if getattr(self, "_cw", None) is not None:
            return self._cw
        w, c = self.w, self.config
        if len(c.dim_mult) != 4 or c.num_res_blocks != 2:
            raise ValueError("b200_wan_vae_decode covers the 4-stage / 3-residual-block decoder of the Wan 2.1 / 2.2 VAE")
        ptr = lambda k: w[k].data_ptr() if k in w else None
        cw = _lib.WanVaeWeights()
        cw.post_quant_w, cw.post_quant_b = ptr("post_quant_conv.weight"), ptr("post_quant_conv.bias")
        cw.conv_in_w, cw.conv_in_b = ptr("decoder.conv_in.weight"), ptr("decoder.conv_in.bias")

        def res(dst, p):
            dst.norm1_gamma, dst.norm2_gamma = ptr(p + ".norm1.gamma"), ptr(p + ".norm2.gamma")
            dst.conv1_w, dst.conv1_b = ptr(p + ".conv1.weight"), ptr(p + ".conv1.bias")
            dst.conv2_w, dst.conv2_b = ptr(p + ".conv2.weight"), ptr(p + ".conv2.bias")
            dst.shortcut_w, dst.shortcut_b = ptr(p + ".conv_shortcut.weight"), ptr(p + ".conv_shortcut.bias")
            dst.cin, dst.cout = w[p + ".norm1.gamma"].numel(), w[p + ".conv1.bias"].numel()

        res(cw.mid_res[0], "decoder.mid_block.resnets.0")
        res(cw.mid_res[1], "decoder.mid_block.resnets.1")
        a = "decoder.mid_block.attentions.0"
        cw.mid_attn.norm_gamma, cw.mid_attn.to_qkv_w, cw.mid_attn.to_qkv_b = ptr(a + ".norm.gamma"), ptr(a + ".to_qkv.weight"), ptr(a + ".to_qkv.bias")
        cw.mid_attn.proj_w, cw.mid_attn.proj_b, cw.mid_attn.channels = ptr(a + ".proj.weight"), ptr(a + ".proj.bias"), self.dims[0]
        for i in range(4):
            for j in range(3):
                res(cw.up_res[i][j], f"decoder.up_blocks.{i}.resnets.{j}")
            if i != 3:
                u = f"decoder.up_blocks.{i}.upsamplers.0"
                cw.up_samp[i].resample_w, cw.up_samp[i].resample_b = ptr(u + ".resample.1.weight"), ptr(u + ".resample.1.bias")
                cw.up_samp[i].time_conv_w, cw.up_samp[i].time_conv_b = ptr(u + ".time_conv.weight"), ptr(u + ".time_conv.bias")
                cw.up_samp[i].channels = w[u + ".resample.1.bias"].numel() * 2
                cw.up_samp[i].temporal = int(bool(self.temporal_upsample[i]))
        cw.norm_out_gamma = ptr("decoder.norm_out.gamma")
        cw.conv_out_w, cw.conv_out_b = ptr("decoder.conv_out.weight"), ptr("decoder.conv_out.bias")
        for i, d in enumerate(self.dims):
            cw.dims[i] = d
        cw.z_pad = self.z_pad
        self._cw = cw
        return cw

    def decode_tile(self, z: torch.Tensor) -> torch.Tensor:
        """z [zc, T, h, w] (one latent tile, all frames) -> planar bf16 [3, 1+4(T-1), 8h, 8w] (pre-clamp) through ONE C call,
        ``b200_wan_vae_decode`` (the whole launch sequence of the decoder is issued from C into a reusable workspace);
        bit-identical to ``decode_tile_py``, which issues the same kernels one by one."""
        import ctypes

        zc, T, h, wd = z.shape
        x = torch.zeros(T, h, wd, 64, dtype=torch.bfloat16, device=z.device)               # K padded to one tile
        x[..., :zc] = z.permute(1, 2, 3, 0)
        cw = self._c_weights()
        lib = _lib.load()
        need = lib.b200_wan_vae_decode_workspace(ctypes.byref(cw), T, h, wd)
        if need < 0:
            _lib.check(int(need), "b200_wan_vae_decode_workspace")
        ws = getattr(self, "_tile_ws", None)
        if ws is None or ws.numel() < need or ws.device != z.device:
            self._tile_ws = ws = torch.empty(int(need), dtype=torch.uint8, device=z.device)
        T_out = 1 + self.config.scale_factor_temporal * (T - 1) if T > 1 else 1
        r = self.spatial_compression_ratio
        out = torch.empty(self.config.out_channels, T_out, h * r, wd * r, dtype=torch.bfloat16, device=z.device)
        rc = lib.b200_wan_vae_decode(x.data_ptr(), ctypes.byref(cw), out.data_ptr(), ws.data_ptr(), ws.numel(), T, h, wd, _stream())
        _lib.check(rc, "b200_wan_vae_decode")
        ops._count(self._tile_launches(T))
        return out

    def _tile_launches(self, T: int) -> int:
        """kernels b200_wan_vae_decode enqueues for a T-frame tile: post_quant + conv_in + conv_out + norm_out (4), the residual
        blocks (norm1, conv1, norm2, conv2; 3 where conv1 + norm2 are one kernel; +1 per shortcut), the mid attention (norm,
        q|k, V^T, output projection + 3 per frame), 3 upsamplers x 2 (+1 time_conv each when T > 1)"""
        w = self.w
        res = [k for k in w if k.endswith(".conv1.bias") and ".resnets." in k]
        n = 4 + sum(3 if conv_norm_fusable(w[k].numel()) else 4 for k in res)
        n += sum(1 for k in w if k.endswith("conv_shortcut.weight")) + 4 + 3 * T + 3 * 2
        if T > 1:
            n += sum(1 for t in self.temporal_upsample if t)
        return n

    def decode_tile_py(self, z: torch.Tensor) -> torch.Tensor:
        """The same decode issued kernel by kernel from Python (round-1 path; kept as the differential partner of the C entry)."""
        w = self.w
        zc, T, h, wd = z.shape
        x = torch.zeros(T, h, wd, 64, dtype=torch.bfloat16, device=z.device)               # K padded to one tile
        x[..., :zc] = z.permute(1, 2, 3, 0)
        x = ops.linear(x.view(-1, 64), w["post_quant_conv.weight"], w["post_quant_conv.bias"]).view(T, h, wd, self.z_pad)
        x = conv3d_cl(x, w["decoder.conv_in.weight"], w["decoder.conv_in.bias"], (3, 3, 3), self.dims[0])
        x = self._res_block(x, "decoder.mid_block.resnets.0")
        x = self._attn_block(x, "decoder.mid_block.attentions.0")
        x = self._res_block(x, "decoder.mid_block.resnets.1")
        n_up = len(self.config.dim_mult)
        for i in range(n_up):
            for j in range(self.config.num_res_blocks + 1):
                x = self._res_block(x, f"decoder.up_blocks.{i}.resnets.{j}")
            if i != n_up - 1:
                x = self._upsample(x, f"decoder.up_blocks.{i}.upsamplers.0", self.temporal_upsample[i])
        x = rmsnorm_silu_cl(x, w["decoder.norm_out.gamma"], out=x)
        return conv3d_cl(x, w["decoder.conv_out.weight"], w["decoder.conv_out.bias"], (3, 3, 3), 16,
                         planar_channels=self.config.out_channels)

    # -------------------------------------------------------------------------------- decode
    @torch.inference_mode()
    def decode(self, z: torch.Tensor, return_dict: bool = False, parallel: Optional[ParallelContext] = None):
        """z [B, zc, T, H, W] -> [B, 3, 1+4(T-1), 8H, 8W] bf16 in [-1, 1]."""
        if not self.w:
            raise RuntimeError("weights not loaded: call load_state_dict()")
        outs = [self._decode_one(z[b].to(self.device), parallel) for b in range(z.shape[0])]
        out = torch.stack(outs, dim=0)
        if return_dict:
            return {"sample": out}
        return (out,)

    def _decode_one(self, z: torch.Tensor, par: Optional[ParallelContext]) -> torch.Tensor:
        zc, T, H, W = z.shape
        r = self.spatial_compression_ratio
        tmin_h, tmin_w = self.tile_sample_min_height // r, self.tile_sample_min_width // r
        if not (self.use_tiling and (W > tmin_w or H > tmin_h)):
            return torch.clamp(self.decode_tile(z), min=-1.0, max=1.0)
        return self.tiled_decode(z, par)

    def tile_grid(self, H: int, W: int) -> List[Tuple[int, int]]:
        r = self.spatial_compression_ratio
        sh, sw = self.tile_sample_stride_height // r, self.tile_sample_stride_width // r
        return [(i, j) for i in range(0, H, sh) for j in range(0, W, sw)]

    def tiled_decode(self, z: torch.Tensor, par: Optional[ParallelContext] = None) -> torch.Tensor:
        """vae/wan/model.py:1516-1623: 32x32-latent tiles at stride 24, in-place blends in row-major tile order,
        crop to the stride, clamp.  Tiles are dealt round-robin over the ranks of ``par`` and all-gathered once."""
        zc, T, H, W = z.shape
        r = self.spatial_compression_ratio
        tmin_h, tmin_w = self.tile_sample_min_height // r, self.tile_sample_min_width // r
        sh, sw = self.tile_sample_stride_height, self.tile_sample_stride_width
        blend_h, blend_w = self.tile_sample_min_height - sh, self.tile_sample_min_width - sw
        if blend_h != blend_w or sh != sw:
            raise ValueError("non-square tiling parameters are not implemented")
        grid = self.tile_grid(H, W)
        ncols = len(range(0, W, sw // r))
        world = par.world_size if par is not None else 1
        rank = par.rank if par is not None else 0
        # tiles are dealt by cost (latent pixels; the edge tiles of a 90 x 160 latent are up to 3.5 x smaller than a full one)
        owners = deal_lpt([min(tmin_h, H - i) * min(tmin_w, W - j) for i, j in grid], world)
        mine = owners[rank]
        T_out = 1 + self.config.scale_factor_temporal * (T - 1)
        tiles: Dict[int, torch.Tensor] = {}
        for idx in mine:
            i, j = grid[idx]
            tiles[idx] = self.decode_tile(z[:, :, i:i + tmin_h, j:j + tmin_w].contiguous())
        if world > 1:
            tiles = self._allgather_tiles(tiles, grid, H, W, T_out, par, owners)
        frame = torch.empty(self.config.out_channels, T_out, H * r, W * r, dtype=torch.bfloat16, device=z.device)
        for idx, (i, j) in enumerate(grid):
            ri, cj = idx // ncols, idx % ncols
            up = tiles[idx - ncols] if ri > 0 else None
            left = tiles[idx - 1] if cj > 0 else None
            blend_tile(tiles[idx], up, left, frame, blend_h, sh, ri * sh, cj * sw)
        return frame

    def _allgather_tiles(self, tiles, grid, H, W, T_out, par: ParallelContext, owners):
        """Pad every rank's tiles to the full tile size, ONE all-gather, slice back to the true tile shapes."""
        r = self.spatial_compression_ratio
        th, tw = self.tile_sample_min_height, self.tile_sample_min_width
        per_rank = max(len(o) for o in owners)
        where = {idx: (src, slot) for src, o in enumerate(owners) for slot, idx in enumerate(o)}
        buf = torch.zeros(per_rank, self.config.out_channels, T_out, th, tw, dtype=torch.bfloat16, device=self.device)
        for slot, idx in enumerate(sorted(tiles)):
            t = tiles[idx]
            buf[slot, :, :, :t.shape[-2], :t.shape[-1]] = t
        allbuf = par.allgather_frames(buf)                                   # [world, per_rank, 3, T, th, tw]
        out = {}
        for idx, (i, j) in enumerate(grid):
            src, slot = where[idx]
            hh = min(th, (H - i) * r)
            ww = min(tw, (W - j) * r)
            out[idx] = allbuf[src, slot, :, :, :hh, :ww].contiguous()
        return out
