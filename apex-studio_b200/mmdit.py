"""Dual-stream (MMDiT) transformer block on the B200 kernels, shared by the HunyuanVideo-1.5 and QwenImage host mirrors
(the Flux mirror in flux/model.py predates this helper and keeps its own copy of the sequence).

The three reference blocks -- ``HunyuanVideo15TransformerBlock.forward`` (transformer/hunyuanvideo15/base/model.py:617-694),
``QwenImageTransformerBlock.forward`` (transformer/qwenimage/base/model.py:679-750), ``FluxTransformerBlock.forward``
(transformer/flux/base/model.py:257-328) -- are the same computation on two token streams that meet in one joint attention:

    per stream:  LayerNorm (no affine) * (1 + scale_msa) + shift_msa  ->  fused q|k|v projection  ->  per-head RMS-norm (+ RoPE)
    joint:       attention over the concatenated sequence (no mask)
    per stream:  h += gate_msa * out_proj(attn);  LayerNorm * (1 + scale_mlp) + shift_mlp;  h += gate_mlp * FF_gelu_tanh(.)

They differ in the order of the streams inside the joint sequence, in which streams are rotated, and in the rounding points
of the per-head norm (``norm_mode`` of ``ops.headnorm_rope_``).  Here ONE residual stream ``h [S_a + S_b, dim]`` holds both
streams as row ranges in the reference's concatenation order, so no concat / split copy exists: the fused QKV GEMMs write
the row ranges of one ``[S, 3*dim]`` buffer, the attention kernel reads q|k|v as strided column blocks through TMA and
writes ``[S, dim]``, and both gate * y + residual updates are GEMM epilogues (12 GEMM-class + 5 row-kernel launches).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import ops


class JointWorkspace:
    """Activation buffers of one forward, shared by all blocks (token-major bf16)."""

    def __init__(self, tokens: int, dim: int, ffn_dim: int, device, ffn_act_dim: int = 0):
        bf = torch.bfloat16
        self.tokens, self.dim, self.ffn_dim = tokens, dim, ffn_dim
        # SwiGLU feed-forwards (Flux2): ffn holds the fused [gate | value] projection, ffn_act the activated product
        self.ffn_act = torch.empty(tokens, ffn_act_dim, dtype=bf, device=device) if ffn_act_dim else None
        self.h = torch.empty(tokens, dim, dtype=bf, device=device)       # residual stream, both streams
        self.norm = torch.empty(tokens, dim, dtype=bf, device=device)
        self.qkv = torch.empty(tokens, 3 * dim, dtype=bf, device=device)
        self.attn = torch.empty(tokens, dim, dtype=bf, device=device)
        self.ffn = torch.empty(tokens, ffn_dim, dtype=bf, device=device)


@dataclass
class StreamParams:
    """One stream of one block: its row range in the joint buffers, the six modulation vectors in the reference's chunk
    order (shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp), weight-key prefixes and its RoPE rows."""
    rows: slice
    mod: Sequence[torch.Tensor]
    qkv: str
    norm_q: str
    norm_k: str
    out: str
    ff: str
    rope: Optional[torch.Tensor]
    swiglu: bool = False      # Flux2FeedForward: linear_in -> silu(x1) * x2 -> linear_out instead of the gelu-tanh FeedForward


def dual_stream_block(w: Dict[str, torch.Tensor], ws: JointWorkspace, streams: Sequence[StreamParams], heads: int,
                      norm_mode: int, eps: float = 1e-6, par=None, n_img_total: int = 0) -> None:
    """One dual-stream block in place on ``ws.h``; see the module docstring.

    Sequence parallel (``par.sp_size > 1``): ``streams[0]`` (the image / latent stream) holds only this rank's token shard
    (``n_img_total`` tokens over the whole group), ``streams[1]`` (text) is replicated; everything but the joint attention
    is token-local, and the attention runs Ulysses-style on this rank's heads over ALL tokens
    (``ParallelContext.joint_tokens_to_heads`` / ``joint_heads_to_tokens``).  Exact -- no windowing."""
    d = ws.dim
    for s in streams:
        shift_msa, scale_msa = s.mod[0], s.mod[1]
        ops.adaln_zero_modulate(ws.h[s.rows], scale_msa, shift_msa, eps=eps, out=ws.norm[s.rows])
        ops.linear(ws.norm[s.rows], w[s.qkv + ".weight"], w.get(s.qkv + ".bias"), out=ws.qkv[s.rows])
        ops.headnorm_rope_(ws.qkv[s.rows, :d], ws.qkv[s.rows, d:2 * d], w[s.norm_q], w[s.norm_k], s.rope, heads, eps, norm_mode)
    attn = ws.attn
    if par is not None and par.sp_size > 1 and par.use_p2p:
        # Ulysses exchange fused into the kernels over NVLink peer memory (parallel.JointPeerExchange).  Buffer reuse across
        # blocks is ordered by the two barriers exactly as in the Wan path (wan/model.py::block): no rank scatters block
        # L+1's q/k/v into a plane before every rank has finished attention L (barrier 1), and no attention L+1 stores into
        # an `o` buffer before its owner has issued the out-projections of block L (they precede its barrier 0 in stream order).
        img, txt = streams[0], streams[1]
        img_first = (img.rows.start or 0) == 0
        n_txt = ws.qkv[txt.rows].shape[0]
        ex = par.joint_peer_exchange(n_img_total, n_txt, heads, 128, img_first, ws.qkv.device)
        for which in range(3):
            ops.rmsnorm_rope_scatter(ws.qkv[img.rows, which * d:(which + 1) * d], None, None, heads, eps, ex.qkv_peers, ex.P,
                                     which * ex.plane + ex.img_off * ex.width, ex.row0, norm=False)
            ex.qkv[which, ex.txt_off:ex.txt_off + n_txt].copy_(
                ws.qkv[txt.rows, which * d + ex.group_col0:which * d + ex.group_col0 + ex.width])
        ex.barrier(0)
        hp = heads // par.sp_size
        as4 = lambda t: t.view(1, ex.S, hp, 128).transpose(1, 2)
        ops.attention_scatter(as4(ex.qkv[0]), as4(ex.qkv[1]), as4(ex.qkv[2]), ex.o_peers, ex.P, ex.n_local, ex.head_off, d,
                              rep_rows=n_txt, rep_first=not img_first)
        ex.barrier(1)
        attn = ex.o      # [n_txt + n_img_local, d] in the local joint order == the row ranges of ws.attn
    elif par is not None and par.sp_size > 1:
        img, txt = streams[0], streams[1]
        img_first = (img.rows.start or 0) == 0
        qkv_h = par.joint_tokens_to_heads(ws.qkv, img.rows, txt.rows, heads, 128, img_first)   # [3, S_joint, (H/P)*128]
        Sj, hp = qkv_h.shape[1], heads // par.sp_size
        o_h = torch.empty_like(qkv_h[0])
        as4 = lambda t: t.view(1, Sj, hp, 128).transpose(1, 2)
        ops.attention(as4(qkv_h[0]), as4(qkv_h[1]), as4(qkv_h[2]), out=as4(o_h))
        par.joint_heads_to_tokens(o_h, n_img_total, img_first, ws.attn, img.rows, txt.rows)
    else:
        S = ws.tokens
        as4 = lambda t: t.view(1, S, heads, 128).transpose(1, 2)
        ops.attention(as4(ws.qkv[:, :d]), as4(ws.qkv[:, d:2 * d]), as4(ws.qkv[:, 2 * d:]), out=as4(ws.attn))
    for s in streams:
        gate_msa, shift_mlp, scale_mlp, gate_mlp = s.mod[2], s.mod[3], s.mod[4], s.mod[5]
        h = ws.h[s.rows]
        ops.linear(attn[s.rows], w[s.out + ".weight"], w.get(s.out + ".bias"), epilogue=ops.EPI_GATE_RES, out=h,
                   gate=gate_msa)
        ops.adaln_zero_modulate(h, scale_mlp, shift_mlp, eps=eps, out=ws.norm[s.rows])
        if s.swiglu:
            ops.linear(ws.norm[s.rows], w[s.ff + ".linear_in.weight"], w.get(s.ff + ".linear_in.bias"), out=ws.ffn[s.rows])
            ops.swiglu(ws.ffn[s.rows], out=ws.ffn_act[s.rows])
            ops.linear(ws.ffn_act[s.rows], w[s.ff + ".linear_out.weight"], w.get(s.ff + ".linear_out.bias"),
                       epilogue=ops.EPI_GATE_RES, out=h, gate=gate_mlp)
        else:
            ops.mlp_gelu_(h, ws.norm[s.rows], w[s.ff + ".net.0.proj.weight"], w.get(s.ff + ".net.0.proj.bias"),
                          w[s.ff + ".net.2.weight"], w.get(s.ff + ".net.2.bias"), gate_mlp, ws.ffn)


def fuse_linears(w: Dict[str, torch.Tensor], dst: str, srcs: Sequence[str]) -> None:
    """Concatenate nn.Linear weights (and biases) row-wise under a new key, removing the sources."""
    w[dst + ".weight"] = torch.cat([w.pop(s + ".weight") for s in srcs], dim=0).contiguous()
    if all((s + ".bias") in w for s in srcs):
        w[dst + ".bias"] = torch.cat([w.pop(s + ".bias") for s in srcs], dim=0).contiguous()


def sinusoid_256(t: torch.Tensor, device) -> torch.Tensor:
    """diffusers Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0): fp32 [cos | sin] of t * 10000^(-i/128)."""
    import math

    half = 128
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=device) / half)
    arg = t[:, None].float() * freqs[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)
