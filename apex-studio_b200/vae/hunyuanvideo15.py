"""HunyuanVideo-1.5 3-D causal VAE -- decode path on the B200 kernels (BASELINE configs[4]).

Host-side mirror of the decode half of the reference's ``AutoencoderKLHunyuanVideo15``
(apps/api/src/vae/hunyuanvideo15/model.py:735; ``decode`` :942-972, ``_decode`` :929-940, ``tiled_decode`` :1060-1119,
``Decoder3D`` :637-732, ``denormalize_latents`` :1145-1150): same ``decoder.*`` state-dict keys and the
``decode(z, return_dict=False)[0]`` / ``enable_tiling`` / ``denormalize_latents`` / ``.dtype`` / ``.config`` surface
``BaseEngine.vae_decode`` and engine/hunyuanvideo15/t2v.py:350-356 rely on.

Design
* activations are channels-last bf16 ``[T, H, W, C]``; every causal conv is the tcgen05 implicit GEMM of csrc/conv.cu.
  This VAE pads with ``mode="replicate"`` (model.py:72-90), which the TMA zero fill of the Wan path cannot express, so the
  producer of every conv input -- channel RMS-norm + SiLU (:366-376, :727-731) or nothing (conv_in, the upsample conv) --
  is ONE gather kernel that writes the replicate-padded tensor (``b200_pad_norm_silu_cl``) and the conv runs on it without
  padding (``b200_conv3d_cl_padded``): no separate pad pass, the norm output never exists unpadded;
* ``HunyuanVideo15Upsample`` = conv -> DCAE channel-to-space rearrangement (first frame not doubled in time) + the
  channel-repeated, rearranged shortcut: ONE kernel after the conv (``b200_dcae_upsample_cl``) instead of the reference's
  2 views, 2 cats, 2 repeat_interleaves and an add;
* mid-block attention is single-head with head dim = channels (1024) over all T*h*w positions of the tile with a
  frame-causal mask (:143-165): scores by the GEMM kernel (fp32 out), ``b200_softmax_rows_block_causal``, P V by the GEMM
  kernel on V^T (row-bias epilogue), output projection with the residual add as epilogue;
* tiling (always on in the engine, t2v.py:350): 8x8-latent tiles at stride 6 (``tile_sample_min`` 128 px, overlap 0.25),
  each decoded for ALL frames; in-place blends in row-major tile order, crop to 96 px, no clamp -- reproduced exactly
  (``b200_blend_tile_noclamp``).  Tiles are independent until the blend: with N GPUs they are dealt round-robin and
  reassembled by ONE all-gather (same scheme as vae/wan.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .. import _lib, ops
from ..parallel import ParallelContext, deal_lpt
from .wan import AutoencoderKLWan, _stream


@dataclass
class HunyuanVideo15VAEConfig:
    """Constructor arguments of the reference class that matter for decode (model.py:747-760)."""
    out_channels: int = 3
    latent_channels: int = 32
    block_out_channels: Sequence[int] = (128, 256, 512, 1024, 1024)     # encoder order, as in the reference config
    layers_per_block: int = 2
    spatial_compression_ratio: int = 16
    temporal_compression_ratio: int = 4
    upsample_match_channel: bool = True
    scaling_factor: float = 1.03682
    shift_factor: Optional[float] = None


# ---------------------------------------------------------------------------------------------------------
# thin wrappers over the C ABI
# ---------------------------------------------------------------------------------------------------------
def pad_norm_silu_cl(x: torch.Tensor, gamma: Optional[torch.Tensor], silu: bool, pads: Tuple[int, int, int] = (2, 1, 1)):
    """x [T,H,W,C] -> replicate-padded [T+pt, H+2ph, W+2pw, C] of (RMS-norm * gamma (+ SiLU))(x), or of x if gamma is None."""
    T, H, W, C = x.shape
    pt, ph, pw = pads
    if not x.is_contiguous():
        raise ValueError("pad_norm_silu_cl needs a contiguous channels-last input")
    out = torch.empty(T + pt, H + 2 * ph, W + 2 * pw, C, dtype=torch.bfloat16, device=x.device)
    rc = _lib.load().b200_pad_norm_silu_cl(x.data_ptr(), out.data_ptr(), None if gamma is None else gamma.data_ptr(), T, H, W, C,
                                           pt, ph, pw, int(silu), _stream())
    _lib.check(rc, "b200_pad_norm_silu_cl")
    ops._count()
    return out


def conv3d_cl_padded(xp: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], cout: int, *,
                     residual: Optional[torch.Tensor] = None, planar_channels: int = 0) -> torch.Tensor:
    """3x3x3 conv on a pre-padded channels-last input [T+2, H+2, W+2, Cin]; w [27*cout, Cin] tap-major."""
    Tp, Hp, Wp, cin = xp.shape
    T, H, W = Tp - 2, Hp - 2, Wp - 2
    if planar_channels:
        out = torch.empty(planar_channels, T, H, W, dtype=torch.bfloat16, device=xp.device)
    else:
        out = torch.empty(T, H, W, cout, dtype=torch.bfloat16, device=xp.device)
    rc = _lib.load().b200_conv3d_cl_padded(xp.data_ptr(), w.data_ptr(), None if bias is None else bias.data_ptr(),
                                           None if residual is None else residual.data_ptr(), out.data_ptr(), T, H, W, cin,
                                           cout, 3, 3, 3, 1 if planar_channels else 0, planar_channels or cout, _stream())
    _lib.check(rc, "b200_conv3d_cl_padded")
    ops._count()
    return out


def dcae_upsample_cl(h: torch.Tensor, x: torch.Tensor, cout: int, temporal: bool) -> torch.Tensor:
    T, H, W, cin = x.shape
    To = 2 * T - 1 if temporal else T
    out = torch.empty(To, 2 * H, 2 * W, cout, dtype=torch.bfloat16, device=x.device)
    rc = _lib.load().b200_dcae_upsample_cl(h.data_ptr(), x.data_ptr(), out.data_ptr(), T, H, W, cout, cin, int(temporal), _stream())
    _lib.check(rc, "b200_dcae_upsample_cl")
    ops._count()
    return out


def softmax_rows_block_causal(s: torch.Tensor, scale: float, block: int) -> torch.Tensor:
    rows, cols = s.shape
    p = torch.empty(rows, cols, dtype=torch.bfloat16, device=s.device)
    rc = _lib.load().b200_softmax_rows_block_causal(s.data_ptr(), p.data_ptr(), rows, cols, s.stride(0), p.stride(0), float(scale),
                                                    block, _stream())
    _lib.check(rc, "b200_softmax_rows_block_causal")
    ops._count()
    return p


def blend_tile_noclamp(tile, up, left, frame, blend: int, crop: int, y0: int, x0: int) -> None:
    planes = tile.shape[0] * tile.shape[1]
    th, tw = tile.shape[-2:]
    uh, uw = (up.shape[-2], up.shape[-1]) if up is not None else (0, 0)
    lh, lw = (left.shape[-2], left.shape[-1]) if left is not None else (0, 0)
    rc = _lib.load().b200_blend_tile_noclamp(tile.data_ptr(), None if up is None else up.data_ptr(),
                                             None if left is None else left.data_ptr(), frame.data_ptr(), planes, th, tw, uh, uw,
                                             lh, lw, blend, min(crop, th), min(crop, tw), y0, x0, frame.shape[-2], frame.shape[-1],
                                             _stream())
    _lib.check(rc, "b200_blend_tile_noclamp")
    ops._count()


# ---------------------------------------------------------------------------------------------------------
class AutoencoderKLHunyuanVideo15:
    """Decode-only B200 implementation (see module docstring)."""

    def __init__(self, config: Optional[HunyuanVideo15VAEConfig] = None, **kwargs):
        self.config = config or HunyuanVideo15VAEConfig(**kwargs)
        c = self.config
        self.dims = list(reversed(list(c.block_out_channels)))          # decoder order
        if not c.upsample_match_channel:
            raise ValueError("upsample_match_channel=False is not implemented")
        if c.latent_channels % 32 or any(d % 32 for d in self.dims) or self.dims[0] % c.latent_channels:
            raise ValueError("channel counts must be multiples of 32 (conv K tiles) and block_out_channels[-1] a multiple of "
                             "latent_channels")
        self.dtype = torch.bfloat16
        self.device = None
        self.w: Dict[str, torch.Tensor] = {}
        self.use_tiling = False
        self.use_light_vae = False
        self.tile_sample_min_height = self.tile_sample_min_width = 128
        self.tile_latent_min_height = self.tile_latent_min_width = 128 // c.spatial_compression_ratio
        self.tile_overlap_factor = 0.25
        self.spatial_compression_ratio = c.spatial_compression_ratio
        self.temporal_compression_ratio = c.temporal_compression_ratio

    # -------------------------------------------------------------------------------- weights
    def _up_flags(self, i: int) -> Tuple[bool, bool]:
        c = self.config
        return i < math.log2(c.spatial_compression_ratio), i < math.log2(c.temporal_compression_ratio)

    def load_state_dict(self, state: Dict[str, torch.Tensor], device="cuda", strict: bool = False):
        """Reference state dict (``decoder.*``; encoder keys are ignored) -> bf16 device tensors in the kernels' layouts:
        3x3x3 conv weights tap-major [27*Cout, Cin], 1x1x1 convs as linear weights, gammas flat; conv_out padded to 16 output
        channels."""
        dev = torch.device(device)
        self.device = dev
        bf = torch.bfloat16
        w: Dict[str, torch.Tensor] = {}
        for k, v in state.items():
            if not k.startswith("decoder."):
                continue
            v = v.detach().to(dev, torch.float32)
            if k == "decoder.conv_out.conv.weight":
                w[k] = AutoencoderKLWan._tap_major(v, cout_pad=16).to(bf)
            elif k == "decoder.conv_out.conv.bias":
                w[k] = torch.nn.functional.pad(v, (0, 16 - v.numel())).to(bf).contiguous()
            elif k.endswith(".weight") and v.dim() == 5 and v.shape[2:] == (1, 1, 1):
                w[k] = v.reshape(v.shape[0], v.shape[1]).to(bf).contiguous()
            elif k.endswith(".weight") and v.dim() == 5:
                w[k] = AutoencoderKLWan._tap_major(v).to(bf)
            elif k.endswith("gamma"):
                w[k] = v.reshape(-1).to(bf).contiguous()
            else:
                w[k] = v.to(bf).contiguous()
        if strict and "decoder.conv_in.conv.weight" not in w:
            raise KeyError("state dict has no decoder.* keys")
        a = "decoder.mid_block.attentions.0"
        if a + ".to_q.weight" in w:    # q|k fused: one GEMM; v stays separate (it is produced transposed)
            w[a + ".to_qk.weight"] = torch.cat([w.pop(a + ".to_q.weight"), w.pop(a + ".to_k.weight")], dim=0).contiguous()
            w[a + ".to_qk.bias"] = torch.cat([w.pop(a + ".to_q.bias"), w.pop(a + ".to_k.bias")], dim=0).contiguous()
        self.w = w
        return [], []

    def init_random_weights(self, device="cuda", seed: int = 7):
        """Synthetic decoder weights in the reference's state-dict format, generated on the device (no checkpoints offline)."""
        dev = torch.device(device)
        g = torch.Generator(device=dev).manual_seed(seed)
        c = self.config
        sd: Dict[str, torch.Tensor] = {}

        def conv(name, cout, cin, k):
            sd[name + ".weight"] = torch.randn(cout, cin, k, k, k, generator=g, device=dev) * (cin * k ** 3) ** -0.5
            sd[name + ".bias"] = torch.randn(cout, generator=g, device=dev) * 0.02

        def gamma(name, ch):
            sd[name] = 1 + torch.randn(ch, generator=g, device=dev) * 0.02

        def res(p, cin, cout):
            gamma(p + ".norm1.gamma", cin)
            conv(p + ".conv1.conv", cout, cin, 3)
            gamma(p + ".norm2.gamma", cout)
            conv(p + ".conv2.conv", cout, cout, 3)
            if cin != cout:
                conv(p + ".conv_shortcut", cout, cin, 1)

        b, p = self.dims, "decoder."
        conv(p + "conv_in.conv", b[0], c.latent_channels, 3)
        res(p + "mid_block.resnets.0", b[0], b[0])
        gamma(p + "mid_block.attentions.0.norm.gamma", b[0])
        for nm in ("to_q", "to_k", "to_v", "proj_out"):
            conv(f"{p}mid_block.attentions.0.{nm}", b[0], b[0], 1)
        res(p + "mid_block.resnets.1", b[0], b[0])
        cur = b[0]
        for i, co in enumerate(b):
            for j in range(c.layers_per_block + 1):
                res(f"{p}up_blocks.{i}.resnets.{j}", cur if j == 0 else co, co)
            cur = co
            sp, tp = self._up_flags(i)
            if sp or tp:
                conv(f"{p}up_blocks.{i}.upsamplers.0.conv.conv", b[i + 1] * (8 if tp else 4), co, 3)
                cur = b[i + 1]
        gamma(p + "norm_out.gamma", b[-1])
        conv(p + "conv_out.conv", c.out_channels, b[-1], 3)
        self.load_state_dict(sd, device=dev)
        return self

    # -------------------------------------------------------------------------------- reference API surface
    def enable_tiling(self, tile_sample_min_height=None, tile_sample_min_width=None, tile_latent_min_height=None,
                      tile_latent_min_width=None, tile_overlap_factor=None, use_light_vae: bool = False) -> None:
        if use_light_vae:
            raise ValueError("the light (TAEHV) decoder is not implemented on the b200 path")
        self.use_tiling = True
        self.tile_sample_min_height = tile_sample_min_height or self.tile_sample_min_height
        self.tile_sample_min_width = tile_sample_min_width or self.tile_sample_min_width
        self.tile_latent_min_height = tile_latent_min_height or self.tile_latent_min_height
        self.tile_latent_min_width = tile_latent_min_width or self.tile_latent_min_width
        self.tile_overlap_factor = tile_overlap_factor or self.tile_overlap_factor

    def disable_tiling(self) -> None:
        self.use_tiling = False

    def denormalize_latents(self, latents: torch.Tensor) -> torch.Tensor:
        c = self.config
        if c.shift_factor:
            return latents / c.scaling_factor + c.shift_factor
        return latents / c.scaling_factor

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    # -------------------------------------------------------------------------------- blocks
    def _resnet(self, x: torch.Tensor, p: str) -> torch.Tensor:
        w = self.w
        cout = w[p + ".conv1.conv.bias"].numel()
        if (p + ".conv_shortcut.weight") in w:
            short = ops.linear(x, w[p + ".conv_shortcut.weight"], w[p + ".conv_shortcut.bias"])
        else:
            short = x
        n = pad_norm_silu_cl(x, w[p + ".norm1.gamma"], True)
        y = conv3d_cl_padded(n, w[p + ".conv1.conv.weight"], w[p + ".conv1.conv.bias"], cout)
        n = pad_norm_silu_cl(y, w[p + ".norm2.gamma"], True)
        return conv3d_cl_padded(n, w[p + ".conv2.conv.weight"], w[p + ".conv2.conv.bias"], cout, residual=short)

    def _attn_block(self, x: torch.Tensor, p: str) -> torch.Tensor:
        """HunyuanVideo15AttnBlock (model.py:167-214): x += proj_out(frame-causal attention over all T*h*w positions)."""
        w = self.w
        T, H, W, C = x.shape
        N = T * H * W
        Np = (N + 7) // 8 * 8          # key axis padded to the GEMM's 16-byte K granularity; pad columns stay zero
        xn = pad_norm_silu_cl(x, w[p + ".norm.gamma"], False, pads=(0, 0, 0)).view(N, C)
        qk = ops.linear(xn, w[p + ".to_qk.weight"], w[p + ".to_qk.bias"])                      # [N, 2C]
        vT = torch.zeros(C, Np, dtype=torch.bfloat16, device=x.device)
        ops.linear(w[p + ".to_v.weight"], xn, w[p + ".to_v.bias"], row_bias=True, out=vT[:, :N])  # [C, N] = V^T
        s = torch.empty(N, Np, dtype=torch.float32, device=x.device)
        ops.linear(qk[:, :C], qk[:, C:], None, epilogue=ops.EPI_BIAS_F32, out=s[:, :N])           # [N, N] fp32 scores
        pm = torch.zeros(N, Np, dtype=torch.bfloat16, device=x.device)
        rc = _lib.load().b200_softmax_rows_block_causal(s.data_ptr(), pm.data_ptr(), N, N, Np, Np, float(C ** -0.5), H * W,
                                                        _stream())
        _lib.check(rc, "b200_softmax_rows_block_causal")
        ops._count()
        o = ops.linear(pm, vT)                                                                   # [N, C]
        ops.linear(o, w[p + ".proj_out.weight"], w[p + ".proj_out.bias"], epilogue=ops.EPI_GATE_RES, out=x.view(N, C), gate=None)
        return x

    def _upsample(self, x: torch.Tensor, p: str, cout: int, temporal: bool) -> torch.Tensor:
        w = self.w
        xp = pad_norm_silu_cl(x, None, False)
        h = conv3d_cl_padded(xp, w[p + ".conv.conv.weight"], w[p + ".conv.conv.bias"], cout * (8 if temporal else 4))
        return dcae_upsample_cl(h, x, cout, temporal)

    def decode_tile(self, z: torch.Tensor) -> torch.Tensor:
        """z [zc, T, h, w] (one latent tile, all frames) -> planar bf16 [3, 1+4(T-1), 16h, 16w]."""
        w, c, b = self.w, self.config, self.dims
        x0 = z.permute(1, 2, 3, 0).contiguous().to(torch.bfloat16)                              # [T,h,w,zc]
        short = x0.repeat_interleave(b[0] // c.latent_channels, dim=-1)                          # model.py:709-711
        x = conv3d_cl_padded(pad_norm_silu_cl(x0, None, False), w["decoder.conv_in.conv.weight"], w["decoder.conv_in.conv.bias"],
                             b[0], residual=short)
        x = self._resnet(x, "decoder.mid_block.resnets.0")
        x = self._attn_block(x, "decoder.mid_block.attentions.0")
        x = self._resnet(x, "decoder.mid_block.resnets.1")
        for i in range(len(b)):
            for j in range(c.layers_per_block + 1):
                x = self._resnet(x, f"decoder.up_blocks.{i}.resnets.{j}")
            sp, tp = self._up_flags(i)
            if sp or tp:
                if not sp:
                    raise ValueError("temporal-only upsampling is not implemented")
                x = self._upsample(x, f"decoder.up_blocks.{i}.upsamplers.0", b[i + 1], tp)
        n = pad_norm_silu_cl(x, w["decoder.norm_out.gamma"], True)
        return conv3d_cl_padded(n, w["decoder.conv_out.conv.weight"], w["decoder.conv_out.conv.bias"], 16,
                                planar_channels=c.out_channels)

    # -------------------------------------------------------------------------------- decode
    @torch.inference_mode()
    def decode(self, z: torch.Tensor, return_dict: bool = False, parallel: Optional[ParallelContext] = None):
        """z [B, zc, T, H, W] -> [B, 3, 1+4(T-1), 16H, 16W] bf16."""
        if not self.w:
            raise RuntimeError("weights not loaded: call load_state_dict()")
        if self.device is None or self.device.type != "cuda":
            raise ValueError("the b200 VAE decode has no CPU fallback: load the weights on a CUDA device")
        outs = [self._decode_one(z[b].to(self.device), parallel) for b in range(z.shape[0])]
        out = torch.stack(outs, dim=0)
        if return_dict:
            return {"sample": out}
        return (out,)

    def _decode_one(self, z: torch.Tensor, par: Optional[ParallelContext]) -> torch.Tensor:
        _, _, H, W = z.shape
        if self.use_tiling and (W > self.tile_latent_min_width or H > self.tile_latent_min_height):
            return self.tiled_decode(z, par)
        return self.decode_tile(z)

    def tile_grid(self, H: int, W: int) -> List[Tuple[int, int]]:
        oh = int(self.tile_latent_min_height * (1 - self.tile_overlap_factor))
        ow = int(self.tile_latent_min_width * (1 - self.tile_overlap_factor))
        return [(i, j) for i in range(0, H, oh) for j in range(0, W, ow)]

    def tiled_decode(self, z: torch.Tensor, par: Optional[ParallelContext] = None) -> torch.Tensor:
        """model.py:1060-1119; tiles dealt round-robin over the ranks of ``par`` and all-gathered once."""
        zc, T, H, W = z.shape
        r = self.spatial_compression_ratio
        if self.tile_sample_min_height != self.tile_sample_min_width or self.tile_latent_min_height != self.tile_latent_min_width:
            raise ValueError("non-square tiling parameters are not implemented")
        tl = self.tile_latent_min_height
        ov = int(tl * (1 - self.tile_overlap_factor))
        blend = int(self.tile_sample_min_height * self.tile_overlap_factor)
        limit = self.tile_sample_min_height - blend
        grid = self.tile_grid(H, W)
        ncols = len(range(0, W, ov))
        nrows = len(range(0, H, ov))
        world = par.world_size if par is not None else 1
        rank = par.rank if par is not None else 0
        owners = deal_lpt([min(tl, H - i) * min(tl, W - j) for i, j in grid], world)     # by cost: edge tiles are smaller
        mine = owners[rank]
        T_out = 1 + self.temporal_compression_ratio * (T - 1)
        tiles: Dict[int, torch.Tensor] = {}
        for idx in mine:
            i, j = grid[idx]
            tiles[idx] = self.decode_tile(z[:, :, i:i + tl, j:j + tl].contiguous())
        if world > 1:
            tiles = self._allgather_tiles(tiles, grid, H, W, T_out, par, owners)
        # output extent: every tile contributes min(limit, its size) pixels (model.py:1113-1117)
        col_w = [min(limit, min(tl, W - j) * r) for j in range(0, W, ov)]
        row_h = [min(limit, min(tl, H - i) * r) for i in range(0, H, ov)]
        frame = torch.empty(self.config.out_channels, T_out, sum(row_h), sum(col_w), dtype=torch.bfloat16, device=z.device)
        y0 = 0
        for ri in range(nrows):
            x0 = 0
            for cj in range(ncols):
                idx = ri * ncols + cj
                up = tiles[idx - ncols] if ri > 0 else None
                left = tiles[idx - 1] if cj > 0 else None
                blend_tile_noclamp(tiles[idx], up, left, frame, blend, limit, y0, x0)
                x0 += col_w[cj]
            y0 += row_h[ri]
        return frame

    def _allgather_tiles(self, tiles, grid, H, W, T_out, par: ParallelContext, owners):
        r = self.spatial_compression_ratio
        th = tw = self.tile_sample_min_height
        tl = self.tile_latent_min_height
        per_rank = max(len(o) for o in owners)
        where = {idx: (src, slot) for src, o in enumerate(owners) for slot, idx in enumerate(o)}
        buf = torch.zeros(per_rank, self.config.out_channels, T_out, th, tw, dtype=torch.bfloat16, device=self.device)
        for slot, idx in enumerate(sorted(tiles)):
            t = tiles[idx]
            buf[slot, :, :, :t.shape[-2], :t.shape[-1]] = t
        allbuf = par.allgather_frames(buf)
        out = {}
        for idx, (i, j) in enumerate(grid):
            src, slot = where[idx]
            hh, ww = min(tl, H - i) * r, min(tl, W - j) * r
            out[idx] = allbuf[src, slot, :, :, :hh, :ww].contiguous()
        return out
