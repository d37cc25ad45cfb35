"""CPU ORACLE (test infrastructure -- never imported by the product package) for the HunyuanVideo-1.5 DiT path
(BASELINE.json configs[4]; SURVEY.md section 8 f1 "dual-stream block variants for HY-1.5").

Functional restatement, in plain torch on the CPU, of ``HunyuanVideo15Transformer3DModel.forward`` (paths relative to
/root/reference/apps/api/src/transformer/hunyuanvideo15/base/model.py):

    :964-1165  forward (embedders, token reorder, 54 dual-stream blocks, head, unpatchify) -> :func:`hy15_forward`
    :617-694   HunyuanVideo15TransformerBlock.forward                                      -> :func:`dual_block`
    :83-171    HunyuanVideo15AttnProcessor2_0                                              -> :func:`joint_attention`
    :450-496   HunyuanVideo15TokenRefiner (+ :297-412 refiner blocks, :205-221 AdaNorm)    -> :func:`token_refiner`
    :224-268   HunyuanVideo15TimeEmbedding                                                 -> :func:`time_embed`
    :499-541   HunyuanVideo15RotaryPosEmbed                                                -> :func:`rope_table`
    :544-582   ByT5 text projection / image projection                                     -> inline
    transformer/efficiency/mod.py:24-35  InplaceRMSNorm (per head here)                    -> wan_dit.rms_norm_across_heads
    transformer/efficiency/ops.py:163-235 apply_cos_sin_rope_inplace                       -> :func:`apply_rope`

Arithmetic of the un-vendored ``diffusers`` dependency (generic Attention + AttnProcessor2_0 of the refiner,
FeedForward "linear-silu" / "gelu-approximate", AdaLayerNormZero / Continuous, Timesteps, TimestepEmbedding,
PixArtAlphaTextProjection, get_1d_rotary_pos_embed) is restated from its published semantics.

fp32 = exact math (with the reference's fp32 aliasing quirk of InplaceRMSNorm reproducible through
``wan_dit.REF_FP32_ALIAS_QUIRK``); bf16 = the reference's rounding points, bit for bit on the CPU.
PINNING: oracle/make_golden.py golden_hy15 runs the reference's own model -> tests/golden/hy15_*.npz;
tests/test_oracle_hy15.py compares (fp32 and bf16).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

import wan_dit
from flux_dit import ada_modulate, feed_forward, linear, sdpa, sinusoid

Weights = Dict[str, torch.Tensor]


def layer_norm(x, w: Weights, p: str, eps: float) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), w[p + ".weight"], w[p + ".bias"], eps)


def timestep_embedder(t_proj: torch.Tensor, w: Weights, p: str) -> torch.Tensor:
    return linear(F.silu(linear(t_proj, w, p + ".linear_1")), w, p + ".linear_2")


def time_embed(timestep: torch.Tensor, w: Weights) -> torch.Tensor:
    """:224-268 -- the sinusoid is cast to the TIMESTEP's dtype (the engine passes it in the latent dtype, t2v.py:243)."""
    return timestep_embedder(sinusoid(timestep).to(timestep.dtype), w, "time_embed.timestep_embedder")


def rope_table(grid: Tuple[int, int, int], rope_dim=(16, 56, 56), theta: float = 256.0):
    """:499-541: meshgrid positions, per axis get_1d_rotary_pos_embed(dim, pos, theta, use_real=True) with its DEFAULT
    float32 frequencies; (cos, sin) [F*H*W, head_dim], every value repeated for its channel pair."""
    axes = [torch.arange(0, n, dtype=torch.float32) for n in grid]
    mesh = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=0)
    cos_out, sin_out = [], []
    for i, d in enumerate(rope_dim):
        freqs = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float32)[: d // 2] / d))
        ang = torch.outer(mesh[i].reshape(-1), freqs)
        cos_out.append(ang.cos().repeat_interleave(2, dim=1).float())
        sin_out.append(ang.sin().repeat_interleave(2, dim=1).float())
    return torch.cat(cos_out, dim=1), torch.cat(sin_out, dim=1)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """efficiency/ops.py:163-235 on x [B,S,H,D]: table cast to x.dtype, every second entry, then (for half precision)
    fp32 complex multiply and one cast back; for fp32 the mul_/addcmul_ pair."""
    c = cos.to(x.dtype)[None, :, None, ::2]
    s = sin.to(x.dtype)[None, :, None, ::2]
    xp = x.unflatten(-1, (-1, 2))
    re, im = xp[..., 0], xp[..., 1]
    if x.dtype in (torch.float16, torch.bfloat16):
        ref, imf, cf, sf = re.float(), im.float(), c.float(), s.float()
        re_o = torch.addcmul(ref * cf, imf, sf, value=-1.0).to(x.dtype)
        im_o = torch.addcmul(imf * cf, ref, sf, value=1.0).to(x.dtype)
    else:
        re_o = torch.addcmul(re * c, im, s, value=-1.0)
        im_o = torch.addcmul(im * c, re, s, value=1.0)
    return torch.stack([re_o, im_o], dim=-1).flatten(-2)


def head_norm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """InplaceRMSNorm(head_dim) on [B,S,H,D] (:603-618)."""
    return wan_dit.rms_norm_across_heads(x, weight, eps)


def joint_attention(w: Weights, p: str, heads: int, x: torch.Tensor, ctx: torch.Tensor, rope):
    """:83-171: latent q/k/v, per-head norm, RoPE on the latent stream only, encoder q/k/v + norm, [latent, encoder]
    concatenation, attention without mask (:1096-1103 passes None on purpose), split, output projections."""
    q = linear(x, w, p + ".to_q").unflatten(2, (heads, -1))
    k = linear(x, w, p + ".to_k").unflatten(2, (heads, -1))
    v = linear(x, w, p + ".to_v").unflatten(2, (heads, -1))
    q, k = head_norm(q, w[p + ".norm_q.weight"]), head_norm(k, w[p + ".norm_k.weight"])
    if rope is not None:
        q, k = apply_rope(q, *rope), apply_rope(k, *rope)
    eq = linear(ctx, w, p + ".add_q_proj").unflatten(2, (heads, -1))
    ek = linear(ctx, w, p + ".add_k_proj").unflatten(2, (heads, -1))
    ev = linear(ctx, w, p + ".add_v_proj").unflatten(2, (heads, -1))
    eq, ek = head_norm(eq, w[p + ".norm_added_q.weight"]), head_norm(ek, w[p + ".norm_added_k.weight"])
    q, k, v = torch.cat([q, eq], dim=1), torch.cat([k, ek], dim=1), torch.cat([v, ev], dim=1)
    o = sdpa(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2).flatten(2, 3).to(q.dtype)
    n = ctx.shape[1]
    return linear(o[:, :-n], w, p + ".to_out.0"), linear(o[:, -n:], w, p + ".to_add_out")


def dual_block(i: int, w: Weights, heads: int, x: torch.Tensor, ctx: torch.Tensor, temb: torch.Tensor, rope):
    """:617-694 -> (x, ctx)."""
    p = f"transformer_blocks.{i}"
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = linear(F.silu(temb), w, p + ".norm1.linear").chunk(6, dim=1)
    c_shift_msa, c_scale_msa, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = linear(
        F.silu(temb), w, p + ".norm1_context.linear").chunk(6, dim=1)
    ax, ac = joint_attention(w, p + ".attn", heads, ada_modulate(x, scale_msa, shift_msa),
                             ada_modulate(ctx, c_scale_msa, c_shift_msa), rope)
    x = x + ax * gate_msa.unsqueeze(1)
    ctx = ctx + ac * c_gate_msa.unsqueeze(1)
    nx = F.layer_norm(x, (x.shape[-1],), None, None, 1e-6) * (1 + scale_mlp[:, None]) + shift_mlp[:, None]
    nc = F.layer_norm(ctx, (ctx.shape[-1],), None, None, 1e-6) * (1 + c_scale_mlp[:, None]) + c_shift_mlp[:, None]
    x = x + gate_mlp.unsqueeze(1) * feed_forward(nx, w, p + ".ff")
    ctx = ctx + c_gate_mlp.unsqueeze(1) * feed_forward(nc, w, p + ".ff_context")
    return x, ctx


def token_refiner(text: torch.Tensor, timestep: torch.Tensor, mask: torch.Tensor, w: Weights, heads: int,
                  num_layers: int) -> torch.Tensor:
    """:450-496 (+ blocks :297-412): masked-mean pooled text -> time/text embedding; proj_in; refiner blocks with a
    key-padding mask (additive -inf [B,1,1,S], :386-402), generic diffusers Attention without q/k norm."""
    p = "context_embedder."
    mf = mask.float().unsqueeze(-1)
    pooled = ((text * mf).sum(dim=1) / mf.sum(dim=1)).to(text.dtype)
    temb = (timestep_embedder(sinusoid(timestep).to(pooled.dtype), w, p + "time_text_embed.timestep_embedder")
            + linear(F.silu(linear(pooled, w, p + "time_text_embed.text_embedder.linear_1")), w,
                     p + "time_text_embed.text_embedder.linear_2"))
    h = linear(text, w, p + "proj_in")
    b, s, _ = h.shape
    attn_mask = None
    if not mask.bool().all():
        attn_mask = torch.zeros(b, 1, 1, s, dtype=h.dtype).masked_fill(~mask.bool().view(b, 1, 1, s), float("-inf"))
    for i in range(num_layers):
        q = p + f"token_refiner.refiner_blocks.{i}"
        n1 = layer_norm(h, w, q + ".norm1", 1e-6)
        qh = linear(n1, w, q + ".attn.to_q").view(b, s, heads, -1).transpose(1, 2)
        kh = linear(n1, w, q + ".attn.to_k").view(b, s, heads, -1).transpose(1, 2)
        vh = linear(n1, w, q + ".attn.to_v").view(b, s, heads, -1).transpose(1, 2)
        o = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=attn_mask, dropout_p=0.0, is_causal=False)
        o = linear(o.transpose(1, 2).reshape(b, s, -1).to(qh.dtype), w, q + ".attn.to_out.0")
        gate_msa, gate_mlp = linear(F.silu(temb), w, q + ".norm_out.linear").chunk(2, dim=1)
        h = h + o * gate_msa.unsqueeze(1)
        n2 = layer_norm(h, w, q + ".norm2", 1e-6)
        ff = linear(F.silu(linear(n2, w, q + ".ff.net.0.proj")), w, q + ".ff.net.2")
        h = h + ff * gate_mlp.unsqueeze(1)
    return h


def condition_tokens(text, mask, text2, mask2, image_embeds, timestep, w: Weights, heads: int, num_refiner_layers: int):
    """:1011-1101: three condition streams + type embedding, then the reorder
    [valid image, valid byt5, valid mllm, invalid image, zeros(invalid byt5), zeros(invalid mllm)]."""
    emb = w["cond_type_embed.weight"]
    t1 = token_refiner(text, timestep, mask, w, heads, num_refiner_layers) + emb[0]
    h2 = layer_norm(text2, w, "context_embedder_2.norm", 1e-5)
    h2 = F.gelu(linear(h2, w, "context_embedder_2.linear_1"))
    h2 = F.gelu(linear(h2, w, "context_embedder_2.linear_2"))
    t2 = linear(h2, w, "context_embedder_2.linear_3") + emb[1]
    h3 = layer_norm(image_embeds, w, "image_embedder.norm_in", 1e-5)
    h3 = linear(F.gelu(linear(h3, w, "image_embedder.linear_1")), w, "image_embedder.linear_2")
    h3 = layer_norm(h3, w, "image_embedder.norm_out", 1e-5)
    is_t2v = bool(torch.all(image_embeds == 0))
    if is_t2v:
        h3 = h3 * 0.0
    mask3 = torch.zeros(h3.shape[:2]) if is_t2v else torch.ones(h3.shape[:2])
    t3 = h3 + emb[2]
    out = []
    for a, ma, b_, mb, c, mc in zip(t1, mask.bool(), t2, mask2.bool(), t3, mask3.bool()):
        out.append(torch.cat([c[mc], b_[mb], a[ma], c[~mc], torch.zeros_like(b_[~mb]), torch.zeros_like(a[~ma])], dim=0))
    return torch.stack(out)


def hy15_forward(hidden: torch.Tensor, timestep: torch.Tensor, text: torch.Tensor, mask: torch.Tensor, text2: torch.Tensor,
                 mask2: torch.Tensor, image_embeds: torch.Tensor, w: Weights, *, heads: int, num_layers: int,
                 num_refiner_layers: int = 2, patch_size: int = 1, patch_size_t: int = 1, rope_dim=(16, 56, 56),
                 rope_theta: float = 256.0) -> torch.Tensor:
    """:964-1165.  hidden [B,C,F,H,W]; timestep [B] in the latent dtype."""
    b, _, f, hh, ww = hidden.shape
    pt, p = patch_size_t, patch_size
    grid = (f // pt, hh // p, ww // p)
    rope = rope_table(grid, rope_dim, rope_theta)
    temb = time_embed(timestep, w)
    x = F.conv3d(hidden, w["x_embedder.proj.weight"], w["x_embedder.proj.bias"], stride=(pt, p, p)).flatten(2).transpose(1, 2)
    ctx = condition_tokens(text, mask, text2, mask2, image_embeds, timestep, w, heads, num_refiner_layers)
    for i in range(num_layers):
        x, ctx = dual_block(i, w, heads, x, ctx, temb, rope)
    scale, shift = linear(F.silu(temb).to(x.dtype), w, "norm_out.linear").chunk(2, dim=1)
    x = F.layer_norm(x, (x.shape[-1],), None, None, 1e-6) * (1 + scale)[:, None, :] + shift[:, None, :]
    x = linear(x, w, "proj_out")
    x = x.reshape(b, grid[0], grid[1], grid[2], -1, pt, p, p).permute(0, 4, 1, 5, 2, 6, 3, 7)
    return x.flatten(6, 7).flatten(4, 5).flatten(2, 3)


# --------------------------------------------------------------------------------------------------
def make_weights(dim: int, heads: int, num_layers: int, num_refiner_layers: int = 2, in_channels: int = 65,
                 out_channels: int = 32, text_dim: int = 3584, text2_dim: int = 1472, image_dim: int = 1152,
                 byt5_hidden: int = 2048, patch_size: int = 1, patch_size_t: int = 1, seed: int = 1234,
                 dtype=torch.float32, std: float = 0.05) -> Weights:
    """Synthetic weights under the reference's state-dict keys (byt5 hidden width is 2048 in the reference, :856-858)."""
    g = torch.Generator().manual_seed(seed)
    hd = dim // heads
    w: Weights = {}

    def lin(name, out_f, in_f, s=std):
        w[name + ".weight"] = torch.randn(out_f, in_f, generator=g) * s
        w[name + ".bias"] = torch.randn(out_f, generator=g) * 0.02

    def ln(name, n):
        w[name + ".weight"] = 1.0 + 0.1 * torch.randn(n, generator=g)
        w[name + ".bias"] = 0.05 * torch.randn(n, generator=g)

    w["x_embedder.proj.weight"] = torch.randn(dim, in_channels, patch_size_t, patch_size, patch_size, generator=g) * std
    w["x_embedder.proj.bias"] = torch.randn(dim, generator=g) * 0.02
    ln("image_embedder.norm_in", image_dim)
    lin("image_embedder.linear_1", image_dim, image_dim)
    lin("image_embedder.linear_2", dim, image_dim)
    ln("image_embedder.norm_out", dim)
    c = "context_embedder."
    lin(c + "time_text_embed.timestep_embedder.linear_1", dim, 256)
    lin(c + "time_text_embed.timestep_embedder.linear_2", dim, dim)
    lin(c + "time_text_embed.text_embedder.linear_1", dim, text_dim)
    lin(c + "time_text_embed.text_embedder.linear_2", dim, dim)
    lin(c + "proj_in", dim, text_dim)
    for i in range(num_refiner_layers):
        q = c + f"token_refiner.refiner_blocks.{i}"
        ln(q + ".norm1", dim)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            lin(f"{q}.attn.{n}", dim, dim)
        ln(q + ".norm2", dim)
        lin(q + ".ff.net.0.proj", 4 * dim, dim)
        lin(q + ".ff.net.2", dim, 4 * dim)
        lin(q + ".norm_out.linear", 2 * dim, dim)
    ln("context_embedder_2.norm", text2_dim)
    lin("context_embedder_2.linear_1", byt5_hidden, text2_dim)
    lin("context_embedder_2.linear_2", byt5_hidden, byt5_hidden, 0.02)
    lin("context_embedder_2.linear_3", dim, byt5_hidden, 0.02)
    lin("time_embed.timestep_embedder.linear_1", dim, 256)
    lin("time_embed.timestep_embedder.linear_2", dim, dim)
    w["cond_type_embed.weight"] = torch.randn(3, dim, generator=g) * 0.5
    for i in range(num_layers):
        p = f"transformer_blocks.{i}"
        lin(p + ".norm1.linear", 6 * dim, dim)
        lin(p + ".norm1_context.linear", 6 * dim, dim)
        for n in ("to_q", "to_k", "to_v", "to_out.0", "add_q_proj", "add_k_proj", "add_v_proj", "to_add_out"):
            lin(f"{p}.attn.{n}", dim, dim)
        for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
            w[f"{p}.attn.{n}.weight"] = 1.0 + 0.1 * torch.randn(hd, generator=g)
        for f_ in ("ff", "ff_context"):
            lin(f"{p}.{f_}.net.0.proj", 4 * dim, dim)
            lin(f"{p}.{f_}.net.2", dim, 4 * dim)
    lin("norm_out.linear", 2 * dim, dim)
    lin("proj_out", patch_size_t * patch_size * patch_size * out_channels, dim, 0.02)
    return {k: v.to(dtype) for k, v in w.items()}
