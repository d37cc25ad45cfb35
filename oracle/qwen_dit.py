"""CPU ORACLE (test infrastructure -- never imported by the product package) for the QwenImage DiT path
(BASELINE.json configs[2]: QwenImage-Edit-2509 1024x1024, 8-step Lightning).

Functional restatement, in plain torch on the CPU, of ``QwenImageTransformer2DModel.forward`` (paths relative to
/root/reference/apps/api/src/transformer/qwenimage/base/model.py):

    :853-993   forward (img_in, txt_norm + txt_in, time embedding, RoPE, 60 dual-stream blocks, head) -> :func:`qwen_forward`
    :679-750   QwenImageTransformerBlock.forward (+ _modulate :639-677 with index=None)                -> :func:`dual_block`
    :480-578   QwenDoubleStreamAttnProcessor2_0                                                        -> :func:`joint_attention`
    :154-183   QwenTimestepProjEmbeddings                                                              -> :func:`time_embed`
    :186-300   QwenEmbedRope (scale_rope=True)                                                         -> :func:`rope_tables`
    :121-151   apply_rotary_emb_qwen(use_real=False)                                                   -> :func:`apply_rope`

Arithmetic of the un-vendored ``diffusers`` dependency (generic Attention container, RMSNorm, FeedForward
"gelu-approximate", AdaLayerNormContinuous, Timesteps(scale=1000), TimestepEmbedding) is restated from its published
semantics.  ``zero_cond_t`` / ``use_additional_t_cond`` / ``use_layer3d_rope`` (later checkpoints) are not restated.
fp32 = exact math; bf16 = the reference's rounding points, bit for bit on the CPU.
PINNING: oracle/make_golden.py golden_qwen -> tests/golden/qwen_*.npz; tests/test_oracle_qwen.py compares.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

from flux_dit import feed_forward, linear, sdpa

Weights = Dict[str, torch.Tensor]


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """diffusers RMSNorm: fp32 statistic; ``x * rsqrt`` promotes to fp32; cast to the (half) weight dtype; gain."""
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    y = x * torch.rsqrt(var + eps)
    if weight.dtype in (torch.float16, torch.bfloat16):
        y = y.to(weight.dtype)
    return y * weight


def time_embed(timestep: torch.Tensor, w: Weights, dtype) -> torch.Tensor:
    """:154-183: Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0, scale=1000) -> TimestepEmbedding."""
    half = 128
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    arg = 1000 * (timestep[:, None].float() * freqs[None, :])
    proj = torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1).to(dtype)
    p = "time_text_embed.timestep_embedder"
    return linear(F.silu(linear(proj, w, p + ".linear_1")), w, p + ".linear_2")


def _rope_params(index: torch.Tensor, dim: int, theta: float = 10000.0) -> torch.Tensor:
    freqs = torch.outer(index, 1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float32).div(dim)))
    return torch.polar(torch.ones_like(freqs), freqs)


def rope_tables(img_shapes: Sequence[Tuple[int, int, int]], txt_len: int, axes_dim=(16, 56, 56), theta: float = 10000.0):
    """QwenEmbedRope.forward with scale_rope=True (:229-300) for ONE sample: ``img_shapes`` lists (frame, height, width) of
    the noisy latent and of every reference image (edit-plus); image idx selects the frame-axis position; height / width
    positions are centred ([-ceil(n/2) .. -1, 0 .. n/2-1]); text positions start at max(height/2, width/2).
    Returns complex64 (img [S_img, 64], txt [txt_len, 64])."""
    pos_index = torch.arange(4096)
    neg_index = torch.arange(4096).flip(0) * -1 - 1
    pos = [_rope_params(pos_index, d, theta) for d in axes_dim]
    neg = [_rope_params(neg_index, d, theta) for d in axes_dim]
    vid, max_vid_index = [], 0
    for idx, (frame, height, width) in enumerate(img_shapes):
        ff = pos[0][idx:idx + frame].view(frame, 1, 1, -1).expand(frame, height, width, -1)
        fh = torch.cat([neg[1][-(height - height // 2):], pos[1][:height // 2]], dim=0).view(1, height, 1, -1).expand(frame, height, width, -1)
        fw = torch.cat([neg[2][-(width - width // 2):], pos[2][:width // 2]], dim=0).view(1, 1, width, -1).expand(frame, height, width, -1)
        vid.append(torch.cat([ff, fh, fw], dim=-1).reshape(frame * height * width, -1))
        max_vid_index = max(height // 2, width // 2, max_vid_index)
    txt = torch.cat(pos, dim=1)[max_vid_index:max_vid_index + txt_len]
    return torch.cat(vid, dim=0), txt


def apply_rope(x: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """:146-151: complex multiply in fp32, one cast back.  x [B,S,H,D], freqs complex [S, D/2]."""
    xc = torch.view_as_complex(x.float().reshape(*x.shape[:-1], -1, 2))
    return torch.view_as_real(xc * freqs.unsqueeze(1)).flatten(3).type_as(x)


def joint_attention(w: Weights, p: str, heads: int, x: torch.Tensor, ctx: torch.Tensor, img_freqs, txt_freqs):
    """:480-578: image and text q/k/v, per-head RMSNorm, RoPE on BOTH streams, [text, image] joint attention."""
    def qkv(t, names):
        return [linear(t, w, f"{p}.{n}").unflatten(-1, (heads, -1)) for n in names]

    iq, ik, iv = qkv(x, ("to_q", "to_k", "to_v"))
    tq, tk, tv = qkv(ctx, ("add_q_proj", "add_k_proj", "add_v_proj"))
    iq, ik = rms_norm(iq, w[p + ".norm_q.weight"]), rms_norm(ik, w[p + ".norm_k.weight"])
    tq, tk = rms_norm(tq, w[p + ".norm_added_q.weight"]), rms_norm(tk, w[p + ".norm_added_k.weight"])
    iq, ik, tq, tk = apply_rope(iq, img_freqs), apply_rope(ik, img_freqs), apply_rope(tq, txt_freqs), apply_rope(tk, txt_freqs)
    q, k, v = torch.cat([tq, iq], dim=1), torch.cat([tk, ik], dim=1), torch.cat([tv, iv], dim=1)
    o = sdpa(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3)).permute(0, 2, 1, 3).flatten(2, 3).to(q.dtype)
    n = ctx.shape[1]
    return linear(o[:, n:], w, p + ".to_out.0"), linear(o[:, :n], w, p + ".to_add_out")


def _modulate(x: torch.Tensor, mod: torch.Tensor):
    shift, scale, gate = mod.chunk(3, dim=-1)
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1), gate.unsqueeze(1)


def dual_block(i: int, w: Weights, heads: int, x: torch.Tensor, ctx: torch.Tensor, temb: torch.Tensor, img_freqs, txt_freqs):
    """:679-750 -> (ctx, x)."""
    p = f"transformer_blocks.{i}"
    ln = lambda t: F.layer_norm(t, (t.shape[-1],), None, None, 1e-6)
    img_mod1, img_mod2 = linear(F.silu(temb), w, p + ".img_mod.1").chunk(2, dim=-1)
    txt_mod1, txt_mod2 = linear(F.silu(temb), w, p + ".txt_mod.1").chunk(2, dim=-1)
    xm, g1 = _modulate(ln(x), img_mod1)
    cm, cg1 = _modulate(ln(ctx), txt_mod1)
    ax, ac = joint_attention(w, p + ".attn", heads, xm, cm, img_freqs, txt_freqs)
    x = x + g1 * ax
    ctx = ctx + cg1 * ac
    xm2, g2 = _modulate(ln(x), img_mod2)
    x = x + g2 * feed_forward(xm2, w, p + ".img_mlp")
    cm2, cg2 = _modulate(ln(ctx), txt_mod2)
    ctx = ctx + cg2 * feed_forward(cm2, w, p + ".txt_mlp")
    return ctx, x


def qwen_forward(hidden: torch.Tensor, enc: torch.Tensor, timestep: torch.Tensor, img_shapes: List, txt_len: int, w: Weights,
                 *, heads: int, num_layers: int, axes_dims_rope=(16, 56, 56)) -> torch.Tensor:
    """:853-993.  hidden [B,S_img,C_in] (noisy latent tokens followed by the reference-image tokens for edit), enc
    [B,S_txt,joint_dim], timestep [B] (sigma, i.e. already / 1000), img_shapes = [(f,h,w), ...] of ONE sample."""
    x = linear(hidden, w, "img_in")
    t = timestep.to(x.dtype)
    ctx = linear(rms_norm(enc, w["txt_norm.weight"]), w, "txt_in")
    temb = time_embed(t, w, x.dtype)
    img_freqs, txt_freqs = rope_tables(img_shapes, txt_len, axes_dims_rope)
    for i in range(num_layers):
        ctx, x = dual_block(i, w, heads, x, ctx, temb, img_freqs, txt_freqs)
    scale, shift = linear(F.silu(temb).to(x.dtype), w, "norm_out.linear").chunk(2, dim=1)
    x = F.layer_norm(x, (x.shape[-1],), None, None, 1e-6) * (1 + scale)[:, None, :] + shift[:, None, :]
    return linear(x, w, "proj_out")


def make_weights(dim: int, heads: int, num_layers: int, in_channels: int = 64, out_channels: int = 16, joint_dim: int = 3584,
                 patch_size: int = 2, seed: int = 1234, dtype=torch.float32, std: float = 0.05) -> Weights:
    g = torch.Generator().manual_seed(seed)
    hd = dim // heads
    w: Weights = {}

    def lin(name, out_f, in_f, s=std):
        w[name + ".weight"] = torch.randn(out_f, in_f, generator=g) * s
        w[name + ".bias"] = torch.randn(out_f, generator=g) * 0.02

    lin("time_text_embed.timestep_embedder.linear_1", dim, 256)
    lin("time_text_embed.timestep_embedder.linear_2", dim, dim)
    w["txt_norm.weight"] = 1.0 + 0.1 * torch.randn(joint_dim, generator=g)
    lin("img_in", dim, in_channels)
    lin("txt_in", dim, joint_dim)
    for i in range(num_layers):
        p = f"transformer_blocks.{i}"
        lin(p + ".img_mod.1", 6 * dim, dim)
        lin(p + ".txt_mod.1", 6 * dim, dim)
        for n in ("to_q", "to_k", "to_v", "to_out.0", "add_q_proj", "add_k_proj", "add_v_proj", "to_add_out"):
            lin(f"{p}.attn.{n}", dim, dim)
        for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
            w[f"{p}.attn.{n}.weight"] = 1.0 + 0.1 * torch.randn(hd, generator=g)
        for f_ in ("img_mlp", "txt_mlp"):
            lin(f"{p}.{f_}.net.0.proj", 4 * dim, dim)
            lin(f"{p}.{f_}.net.2", dim, 4 * dim)
    lin("norm_out.linear", 2 * dim, dim)
    lin("proj_out", patch_size * patch_size * out_channels, dim, 0.02)
    return {k: v.to(dtype) for k, v in w.items()}
