"""CPU ORACLE (test infrastructure -- never imported by the product package) for the Flux2 DiT path
(BASELINE.json configs[0]: Flux2-Klein-4B, the reference's CPU-runnable plumbing case; SURVEY.md section 8 f1 "SwiGLU MLP for Flux2").

Functional restatement, in plain torch on the CPU, of ``Flux2Transformer2DModel.forward`` (paths relative to
/root/reference/apps/api/src/transformer/flux2/base/model.py):

    :886-1032  forward (shared modulation, embedders, RoPE over [text, image], 8 dual + 48 single blocks, head) -> :func:`flux2_forward`
    :521-627   Flux2TransformerBlock                                   -> :func:`dual_block`
    :449-518   Flux2SingleTransformerBlock (+ parallel attention/MLP processor :300-356) -> :func:`single_block`
    :133-203   Flux2AttnProcessor                                      -> :func:`joint_attention`
    :91-130    Flux2SwiGLU / Flux2FeedForward                          -> :func:`feed_forward`
    :630-659   Flux2PosEmbed (4 axes, theta 2000)                      -> flux_dit.rope_table
    :662-725   Flux2TimestepGuidanceEmbeddings / Flux2Modulation       -> :func:`time_guidance_embed`, :func:`modulation`

diffusers pieces (TimestepEmbedding without bias, AdaLayerNormContinuous, get_1d_rotary_pos_embed, apply_rotary_emb) are
restated from their published semantics.  fp32 = exact math; bf16 = the reference's rounding points bit for bit on the CPU.
PINNING: oracle/make_golden.py golden_flux2 -> tests/golden/flux2_*.npz; tests/test_oracle_flux2.py compares."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from flux_dit import apply_rope, head_rms_norm, linear, rope_table, sdpa, sinusoid

Weights = Dict[str, torch.Tensor]


def time_guidance_embed(timestep: torch.Tensor, guidance: Optional[torch.Tensor], w: Weights) -> torch.Tensor:
    """:688-702: the sinusoid is cast to the (bf16) timestep's dtype; TimestepEmbedding without biases."""
    p = "time_guidance_embed."
    emb = linear(F.silu(linear(sinusoid(timestep).to(timestep.dtype), w, p + "timestep_embedder.linear_1")), w,
                 p + "timestep_embedder.linear_2")
    if guidance is not None and (p + "guidance_embedder.linear_1.weight") in w:
        emb = emb + linear(F.silu(linear(sinusoid(guidance).to(guidance.dtype), w, p + "guidance_embedder.linear_1")), w,
                           p + "guidance_embedder.linear_2")
    return emb


def modulation(temb: torch.Tensor, w: Weights, p: str, sets: int):
    """Flux2Modulation (:705-725): SiLU -> Linear -> [B,1,3*sets*dim] chunks, grouped as (shift, scale, gate) triples."""
    mod = linear(F.silu(temb), w, p + ".linear").unsqueeze(1)
    ch = torch.chunk(mod, 3 * sets, dim=-1)
    return tuple(ch[3 * i:3 * (i + 1)] for i in range(sets))


def feed_forward(x: torch.Tensor, w: Weights, p: str) -> torch.Tensor:
    x1, x2 = linear(x, w, p + ".linear_in").chunk(2, dim=-1)
    return linear(F.silu(x1) * x2, w, p + ".linear_out")


def joint_attention(w: Weights, p: str, heads: int, x: torch.Tensor, ctx: torch.Tensor, rope):
    """:133-203: [text, image] joint attention, torch.nn.RMSNorm per head, RoPE over the joint sequence."""
    split = lambda t, n: linear(t, w, f"{p}.{n}").unflatten(-1, (heads, -1))
    q, k, v = split(x, "to_q"), split(x, "to_k"), split(x, "to_v")
    q, k = head_rms_norm(q, w[p + ".norm_q.weight"]), head_rms_norm(k, w[p + ".norm_k.weight"])
    eq, ek, ev = split(ctx, "add_q_proj"), split(ctx, "add_k_proj"), split(ctx, "add_v_proj")
    eq, ek = head_rms_norm(eq, w[p + ".norm_added_q.weight"]), head_rms_norm(ek, w[p + ".norm_added_k.weight"])
    q, k, v = torch.cat([eq, q], dim=1), torch.cat([ek, k], dim=1), torch.cat([ev, v], dim=1)
    q, k = apply_rope(q, *rope), apply_rope(k, *rope)
    o = sdpa(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2).flatten(2, 3).to(q.dtype)
    n = ctx.shape[1]
    return linear(o[:, n:], w, p + ".to_out.0"), linear(o[:, :n], w, p + ".to_add_out")


def dual_block(i: int, w: Weights, heads: int, x: torch.Tensor, ctx: torch.Tensor, mod_img, mod_txt, rope):
    """:558-627 -> (ctx, x)."""
    p = f"transformer_blocks.{i}"
    ln = lambda t: F.layer_norm(t, (t.shape[-1],), None, None, 1e-6)
    (shift_msa, scale_msa, gate_msa), (shift_mlp, scale_mlp, gate_mlp) = mod_img
    (c_shift_msa, c_scale_msa, c_gate_msa), (c_shift_mlp, c_scale_mlp, c_gate_mlp) = mod_txt
    nx = (1 + scale_msa) * ln(x) + shift_msa
    nc = (1 + c_scale_msa) * ln(ctx) + c_shift_msa
    ax, ac = joint_attention(w, p + ".attn", heads, nx, nc, rope)
    x = x + gate_msa * ax
    x = x + gate_mlp * feed_forward(ln(x) * (1 + scale_mlp) + shift_mlp, w, p + ".ff")
    ctx = ctx + c_gate_msa * ac
    ctx = ctx + c_gate_mlp * feed_forward(ln(ctx) * (1 + c_scale_mlp) + c_shift_mlp, w, p + ".ff_context")
    return ctx, x


def single_block(i: int, w: Weights, heads: int, h: torch.Tensor, mod, rope) -> torch.Tensor:
    """:479-518 with the parallel attention + MLP processor (:311-356) on the concatenated [text, image] stream."""
    p = f"single_transformer_blocks.{i}.attn"
    shift, scale, gate = mod
    d = h.shape[-1]
    n = (1 + scale) * F.layer_norm(h, (d,), None, None, 1e-6) + shift
    proj = linear(n, w, p + ".to_qkv_mlp_proj")
    qkv, mlp = proj[..., :3 * d], proj[..., 3 * d:]
    q, k, v = (t.unflatten(-1, (heads, -1)) for t in qkv.chunk(3, dim=-1))
    q, k = head_rms_norm(q, w[p + ".norm_q.weight"]), head_rms_norm(k, w[p + ".norm_k.weight"])
    q, k = apply_rope(q, *rope), apply_rope(k, *rope)
    o = sdpa(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2).flatten(2, 3).to(q.dtype)
    m1, m2 = mlp.chunk(2, dim=-1)
    out = linear(torch.cat([o, F.silu(m1) * m2], dim=-1), w, p + ".to_out")
    return h + gate * out


def flux2_forward(hidden: torch.Tensor, enc: torch.Tensor, timestep: torch.Tensor, img_ids: torch.Tensor, txt_ids: torch.Tensor,
                  guidance: Optional[torch.Tensor], w: Weights, *, heads: int, num_layers: int, num_single_layers: int,
                  axes_dims_rope=(32, 32, 32, 32), rope_theta: float = 2000.0) -> torch.Tensor:
    """:886-1032.  timestep [B] is sigma (already / 1000)."""
    n_txt = enc.shape[1]
    t = timestep.to(hidden.dtype) * 1000
    g = guidance.to(hidden.dtype) * 1000 if guidance is not None else None
    temb = time_guidance_embed(t, g, w)
    mod_img = modulation(temb, w, "double_stream_modulation_img", 2)
    mod_txt = modulation(temb, w, "double_stream_modulation_txt", 2)
    mod_single = modulation(temb, w, "single_stream_modulation", 1)[0]
    x, ctx = linear(hidden, w, "x_embedder"), linear(enc, w, "context_embedder")
    ri, rt = rope_table(img_ids, axes_dims_rope, rope_theta), rope_table(txt_ids, axes_dims_rope, rope_theta)
    rope = (torch.cat([rt[0], ri[0]], dim=0), torch.cat([rt[1], ri[1]], dim=0))
    for i in range(num_layers):
        ctx, x = dual_block(i, w, heads, x, ctx, mod_img, mod_txt, rope)
    h = torch.cat([ctx, x], dim=1)
    for i in range(num_single_layers):
        h = single_block(i, w, heads, h, mod_single, rope)
    x = h[:, n_txt:]
    scale, shift = linear(F.silu(temb).to(x.dtype), w, "norm_out.linear").chunk(2, dim=1)
    x = F.layer_norm(x, (x.shape[-1],), None, None, 1e-6) * (1 + scale)[:, None, :] + shift[:, None, :]
    return linear(x, w, "proj_out")


def make_weights(dim: int, heads: int, num_layers: int, num_single_layers: int, in_channels: int = 128, joint_dim: int = 15360,
                 mlp_ratio: float = 3.0, guidance_embeds: bool = False, seed: int = 1234, dtype=torch.float32,
                 std: float = 0.05) -> Weights:
    """Synthetic weights under the reference's state-dict keys (no biases anywhere, model.py:791-884)."""
    g = torch.Generator().manual_seed(seed)
    hd, mlp = dim // heads, int(dim * mlp_ratio)
    w: Weights = {}

    def lin(name, out_f, in_f, s=std):
        w[name + ".weight"] = torch.randn(out_f, in_f, generator=g) * s

    for e in ["timestep_embedder"] + (["guidance_embedder"] if guidance_embeds else []):
        lin(f"time_guidance_embed.{e}.linear_1", dim, 256)
        lin(f"time_guidance_embed.{e}.linear_2", dim, dim)
    lin("double_stream_modulation_img.linear", 6 * dim, dim)
    lin("double_stream_modulation_txt.linear", 6 * dim, dim)
    lin("single_stream_modulation.linear", 3 * dim, dim)
    lin("x_embedder", dim, in_channels)
    lin("context_embedder", dim, joint_dim)
    for i in range(num_layers):
        p = f"transformer_blocks.{i}"
        for n in ("to_q", "to_k", "to_v", "to_out.0", "add_q_proj", "add_k_proj", "add_v_proj", "to_add_out"):
            lin(f"{p}.attn.{n}", dim, dim)
        for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
            w[f"{p}.attn.{n}.weight"] = 1.0 + 0.1 * torch.randn(hd, generator=g)
        for f_ in ("ff", "ff_context"):
            lin(f"{p}.{f_}.linear_in", 2 * mlp, dim)
            lin(f"{p}.{f_}.linear_out", dim, mlp)
    for i in range(num_single_layers):
        p = f"single_transformer_blocks.{i}.attn"
        lin(p + ".to_qkv_mlp_proj", 3 * dim + 2 * mlp, dim)
        w[p + ".norm_q.weight"] = 1.0 + 0.1 * torch.randn(hd, generator=g)
        w[p + ".norm_k.weight"] = 1.0 + 0.1 * torch.randn(hd, generator=g)
        lin(p + ".to_out", dim, dim + mlp)
    lin("norm_out.linear", 2 * dim, dim)
    lin("proj_out", in_channels, dim, 0.02)
    return {k: v.to(dtype) for k, v in w.items()}
