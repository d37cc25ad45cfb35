"""CPU ORACLE (test infrastructure -- never imported by the product package) for the Wan DiT hot path.

A functional restatement, in plain torch on the CPU, of what the reference computes per denoise step
inside ``WanTransformer3DModel.forward`` (all paths relative to /root/reference/apps/api/src/):

    transformer/wan/base/model.py:1684-1891   forward            -> :func:`dit_forward`
    transformer/wan/base/model.py:1101-1333   block forward      -> :func:`block_forward`
    transformer/wan/base/attention.py:305-413 attention processor-> :func:`attention_layer`
    transformer/wan/base/model.py:773-823     time/text embedder -> :func:`condition_embed`
    transformer/wan/base/model.py:826-951     3-axis RoPE table  -> :func:`rope_table`
    transformer/efficiency/mod.py:24-35       InplaceRMSNorm     -> :func:`rms_norm_across_heads`
    transformer/efficiency/ops.py:101-160     RoPE               -> :func:`apply_rope`
    transformer/efficiency/ops.py:19-56       gate / scale-shift -> inline
    attention/functions.py:338-377            `sdpa` backend     -> :func:`sdpa`
    engine/wan/shared/__init__.py:565         CFG combine        -> :func:`cfg_combine`

Arithmetic that lives in the un-vendored, un-pinned dependency ``diffusers`` (FeedForward with
"gelu-approximate", FP32LayerNorm, Timesteps, TimestepEmbedding, PixArtAlphaTextProjection) is restated
from its published semantics (see oracle/ref_import/diffusers/__init__.py).

Weights are a flat ``dict[str, Tensor]`` using the reference's own (diffusers-format) state-dict keys.
Run with every tensor in float32 it is the exact-math oracle; run with bf16 weights/activations every
torch op rounds exactly where the reference rounds (SURVEY.md section 8 "Rounding points"), so on the CPU
it reproduces the reference's bf16 pipeline bit for bit.

PINNING: oracle/make_golden.py runs the reference's own modules (imported from /root/reference through
the stand-in) on seeded inputs and stores input/output vectors under tests/golden/; tests/test_oracle_*.py
check this restatement against them.  The reference itself ships no golden vectors for this path
(SURVEY.md section 4); the only numeric pin of its own is the attention differential recipe
(scripts/smoke_tests/test_attention_backends.py:232-388), which tests/ reproduce.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Weights = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------------
def linear(x: torch.Tensor, w: Weights, prefix: str) -> torch.Tensor:
    return F.linear(x, w[prefix + ".weight"], w.get(prefix + ".bias"))


def fp32_layer_norm(x: torch.Tensor, eps: float, weight=None, bias=None) -> torch.Tensor:
    """diffusers FP32LayerNorm: statistics and affine in fp32, result cast back."""
    return F.layer_norm(
        x.float(), (x.shape[-1],), None if weight is None else weight.float(), None if bias is None else bias.float(), eps
    ).to(x.dtype)


def modulated_norm(x: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, eps: float) -> torch.Tensor:
    """model.py:56-116 with chunk_size=None: norm -> x.addcmul_(x, scale) -> x.add_(shift) (ops.py:37-56)."""
    out = fp32_layer_norm(x, eps)
    out = torch.addcmul(out, out, scale)
    return out + shift


#: The reference's InplaceRMSNorm does ``y = x.float(); y.pow_(2)`` (mod.py:27-28).  For bf16/fp16 inputs
#: ``x.float()`` is a copy and the code computes a plain RMS norm.  For float32 inputs ``x.float()`` IS ``x``,
#: so ``pow_`` squares the projection in place before it is normalised -- the reference's fp32 path is not
#: the exact-math version of its bf16 path.  The production path is bf16 (utils/dtype.py:96-100), so the
#: oracle implements the bf16 semantics; set this flag to reproduce the fp32 quirk bit for bit (used only
#: to pin the oracle against the reference's fp32 golden output).
REF_FP32_ALIAS_QUIRK = False


def rms_norm_across_heads(x: torch.Tensor, weight: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    """mod.py:24-35: fp32 mean of squares over ALL channels, factor cast to x.dtype, two multiplies."""
    if REF_FP32_ALIAS_QUIRK and x.dtype == torch.float32:
        x = x.pow(2)
        y = x.mean(dim=-1, keepdim=True).add(eps).rsqrt()
        x = x * y
        return x * weight if weight is not None else x
    y = x.float().pow(2).mean(dim=-1, keepdim=True).add(eps).rsqrt()
    x = x * y.to(x.dtype)
    if weight is not None:
        x = x * weight.to(x.dtype)
    return x


def rope_1d(dim: int, length: int, theta: float = 10000.0, start: int = 0) -> torch.Tensor:
    """model.py:826-844 (complex128 table [length, dim/2], positions start..start+length-1)."""
    base = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float64) / dim))
    pos = torch.arange(start, start + length, dtype=torch.float64)
    ang = torch.outer(pos, base)
    return torch.polar(torch.ones_like(ang), ang)


def rope_table(head_dim: int, grid: Tuple[int, int, int], max_seq_len: int = 1024, theta: float = 10000.0,
               time_offset: int = -1) -> torch.Tensor:
    """model.py:853-951: head_dim split t|h|w with h = w = 2*(head_dim//6); time rows start at index 1 when
    the table carries the t=-1 sentinel row.  Returns complex128 [F*H*W, head_dim/2]."""
    f, h, w = grid
    h_dim = w_dim = 2 * (head_dim // 6)
    t_dim = head_dim - h_dim - w_dim
    t_len = max_seq_len + (1 if time_offset < 0 else 0)
    ft = rope_1d(t_dim, t_len, theta, start=time_offset)
    fh = rope_1d(h_dim, max_seq_len, theta, 0)
    fw = rope_1d(w_dim, max_seq_len, theta, 0)
    t_start = 1 if time_offset < 0 else 0
    t3 = ft[t_start:t_start + f].view(f, 1, 1, -1).expand(f, h, w, -1)
    h3 = fh[:h].view(1, h, 1, -1).expand(f, h, w, -1)
    w3 = fw[:w].view(1, 1, w, -1).expand(f, h, w, -1)
    return torch.cat([t3, h3, w3], dim=-1).reshape(f * h * w, head_dim // 2)


def apply_rope(x: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """ops.py:101-160.  x [B,H,S,D]; freqs complex [S, D/2].  cos/sin are cast to x.dtype FIRST, then
    re' = re*c - im*s via mul_ + addcmul_ (each an op in x.dtype)."""
    c = freqs.real.to(x.dtype)[None, None]
    s = freqs.imag.to(x.dtype)[None, None]
    xp = x.unflatten(-1, (-1, 2))
    re, im = xp[..., 0], xp[..., 1]
    re_out = torch.addcmul(re * c, im, s, value=-1.0)
    im_out = torch.addcmul(im * c, re, s, value=1.0)
    return torch.stack([re_out, im_out], dim=-1).flatten(-2)


def sdpa(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: Optional[float] = None) -> torch.Tensor:
    """attention/functions.py:338-377 -> F.scaled_dot_product_attention(q, k, v, scale=softmax_scale)."""
    return F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False, scale=scale)


def sdpa_fp32_math(q, k, v, scale: Optional[float] = None) -> torch.Tensor:
    """Exact-math form (fp32 softmax(q k^T * scale) v) used as the tight reference for the CUDA kernel."""
    q, k, v = q.float(), k.float(), v.float()
    scale = scale if scale is not None else 1.0 / math.sqrt(q.shape[-1])
    return torch.softmax((q @ k.transpose(-1, -2)) * scale, dim=-1) @ v


def feed_forward(x: torch.Tensor, w: Weights, prefix: str) -> torch.Tensor:
    """diffusers FeedForward(activation_fn="gelu-approximate"): net.0.proj -> gelu(tanh) -> net.2."""
    h = F.gelu(linear(x, w, prefix + ".net.0.proj"), approximate="tanh")
    return linear(h, w, prefix + ".net.2")


# --------------------------------------------------------------------------------------------------
# attention layer / block / model
# --------------------------------------------------------------------------------------------------
def attention_layer(x: torch.Tensor, w: Weights, prefix: str, heads: int, eps: float,
                    context: Optional[torch.Tensor] = None, freqs: Optional[torch.Tensor] = None) -> torch.Tensor:
    """attention.py:305-413 (no image-context branch, no kv cache): q/k/v Linear, q/k RMS-norm across heads,
    [B,H,S,Dh] views, RoPE on q and k (self-attention only), attention, to_out[0]."""
    ctx = x if context is None else context
    q = linear(x, w, prefix + ".to_q")
    k = linear(ctx, w, prefix + ".to_k")
    v = linear(ctx, w, prefix + ".to_v")
    q = rms_norm_across_heads(q, w.get(prefix + ".norm_q.weight"), eps)
    k = rms_norm_across_heads(k, w.get(prefix + ".norm_k.weight"), eps)
    q = q.unflatten(2, (heads, -1)).transpose(1, 2)
    k = k.unflatten(2, (heads, -1)).transpose(1, 2)
    v = v.unflatten(2, (heads, -1)).transpose(1, 2)
    if freqs is not None:
        q = apply_rope(q, freqs)
        k = apply_rope(k, freqs)
    o = sdpa(q, k, v).transpose(1, 2).flatten(2, 3).type_as(q)
    return linear(o, w, prefix + ".to_out.0")


def block_forward(h: torch.Tensor, ctx: torch.Tensor, temb6: torch.Tensor, freqs: torch.Tensor, w: Weights,
                  prefix: str, heads: int, eps: float = 1e-6, cross_attn_norm: bool = True) -> torch.Tensor:
    """model.py:1101-1333 (inference path, temb6 [B,6,dim])."""
    table = w[prefix + ".scale_shift_table"]
    shift_msa, scale_msa, gate_msa, c_shift, c_scale, c_gate = (table + temb6.float()).to(h.dtype).chunk(6, dim=1)
    # 1. self-attention
    n = modulated_norm(h, scale_msa, shift_msa, eps)
    a = attention_layer(n, w, prefix + ".attn1", heads, eps, None, freqs)
    h = h + a * gate_msa          # apply_gate_inplace (rounds) then hidden_states.add_ (rounds)
    # 2. cross-attention
    if cross_attn_norm:
        n = fp32_layer_norm(h, eps, w[prefix + ".norm2.weight"], w[prefix + ".norm2.bias"])
    else:
        n = h
    a = attention_layer(n, w, prefix + ".attn2", heads, eps, ctx, None)
    h = h + a
    # 3. feed-forward
    n = modulated_norm(h, c_scale, c_shift, eps)
    f = feed_forward(n, w, prefix + ".ffn")
    h = h + f * c_gate
    return h


def timestep_sinusoid(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): fp32 [B, dim] = [cos | sin]."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    arg = t[:, None].float() * freqs[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


def condition_embed(t: torch.Tensor, text: torch.Tensor, w: Weights, freq_dim: int):
    """model.py:773-823.  Returns temb [B,dim], timestep_proj [B,6,dim], text context [B,L,dim]."""
    p = "condition_embedder."
    dt = w[p + "time_embedder.linear_1.weight"].dtype
    ts = timestep_sinusoid(t, freq_dim).to(dt)
    temb = linear(F.silu(linear(ts, w, p + "time_embedder.linear_1")), w, p + "time_embedder.linear_2").type_as(text)
    tproj = linear(F.silu(temb), w, p + "time_proj")
    ctx = linear(F.gelu(linear(text, w, p + "text_embedder.linear_1"), approximate="tanh"), w,
                 p + "text_embedder.linear_2")
    return temb, tproj.unflatten(1, (6, -1)), ctx


def patchify(latents: torch.Tensor, w: Weights, patch=(1, 2, 2)) -> torch.Tensor:
    """model.py:1748-1749: Conv3d(kernel=stride=patch) -> flatten(2).transpose(1,2) -> [B, F*H*W, dim]."""
    x = F.conv3d(latents, w["patch_embedding.weight"], w["patch_embedding.bias"], stride=patch)
    return x.flatten(2).transpose(1, 2)


def unpatchify(x: torch.Tensor, grid: Tuple[int, int, int], patch=(1, 2, 2)) -> torch.Tensor:
    """model.py:1870-1882."""
    b = x.shape[0]
    f, h, w_ = grid
    pt, ph, pw = patch
    x = x.reshape(b, f, h, w_, pt, ph, pw, -1).permute(0, 7, 1, 4, 2, 5, 3, 6)
    return x.flatten(6, 7).flatten(4, 5).flatten(2, 3)


def dit_forward(latents: torch.Tensor, timestep: torch.Tensor, text: torch.Tensor, w: Weights, *, heads: int,
                num_layers: int, freq_dim: int = 256, eps: float = 1e-6, patch=(1, 2, 2),
                rope_max_seq_len: int = 1024, cross_attn_norm: bool = True) -> torch.Tensor:
    """model.py:1684-1891 (t2v: no image context, no ip-adapter).  latents [B,C,F,H,W] -> same shape."""
    b, c, f, hh, ww = latents.shape
    grid = (f // patch[0], hh // patch[1], ww // patch[2])
    dim = w["patch_embedding.weight"].shape[0]
    freqs = rope_table(dim // heads, grid, rope_max_seq_len)
    h = patchify(latents, w, patch)
    temb, temb6, ctx = condition_embed(timestep, text, w, freq_dim)
    for i in range(num_layers):
        h = block_forward(h, ctx, temb6, freqs, w, f"blocks.{i}", heads, eps, cross_attn_norm)
    shift, scale = (w["scale_shift_table"] + temb.unsqueeze(1)).chunk(2, dim=1)
    h = modulated_norm(h, scale, shift, eps)
    h = linear(h, w, "proj_out")
    return unpatchify(h, grid, patch)


def cfg_combine(cond: torch.Tensor, uncond: torch.Tensor, guidance_scale: float) -> torch.Tensor:
    """engine/wan/shared/__init__.py:565."""
    return uncond + guidance_scale * (cond - uncond)


# --------------------------------------------------------------------------------------------------
# synthetic weights: N(0, 0.02^2) linears, scale_shift_table randn/sqrt(dim),
# norm weights 1 + N(0, 0.02^2); generator seeded on the CPU.
# --------------------------------------------------------------------------------------------------
def make_weights(*, dim: int, heads: int, ffn_dim: int, num_layers: int, in_channels: int = 16, out_channels: int = 16,
                 text_dim: int = 4096, freq_dim: int = 256, patch=(1, 2, 2), seed: int = 1234,
                 dtype=torch.float32, std: float = 0.02) -> Weights:
    g = torch.Generator(device="cpu").manual_seed(seed)
    w: Weights = {}

    def lin(name, out_f, in_f):
        w[name + ".weight"] = torch.randn(out_f, in_f, generator=g) * std
        w[name + ".bias"] = torch.randn(out_f, generator=g) * std

    pp = patch[0] * patch[1] * patch[2]
    w["patch_embedding.weight"] = (torch.randn(dim, in_channels * pp, generator=g) * std).view(dim, in_channels, *patch)
    w["patch_embedding.bias"] = torch.randn(dim, generator=g) * std
    lin("condition_embedder.time_embedder.linear_1", dim, freq_dim)
    lin("condition_embedder.time_embedder.linear_2", dim, dim)
    lin("condition_embedder.time_proj", dim * 6, dim)
    lin("condition_embedder.text_embedder.linear_1", dim, text_dim)
    lin("condition_embedder.text_embedder.linear_2", dim, dim)
    for i in range(num_layers):
        p = f"blocks.{i}"
        w[p + ".scale_shift_table"] = torch.randn(1, 6, dim, generator=g) / dim ** 0.5
        for a in ("attn1", "attn2"):
            for n in ("to_q", "to_k", "to_v", "to_out.0"):
                lin(f"{p}.{a}.{n}", dim, dim)
            w[f"{p}.{a}.norm_q.weight"] = 1 + torch.randn(dim, generator=g) * std
            w[f"{p}.{a}.norm_k.weight"] = 1 + torch.randn(dim, generator=g) * std
        w[p + ".norm2.weight"] = 1 + torch.randn(dim, generator=g) * std
        w[p + ".norm2.bias"] = torch.randn(dim, generator=g) * std
        lin(p + ".ffn.net.0.proj", ffn_dim, dim)
        lin(p + ".ffn.net.2", dim, ffn_dim)
    w["scale_shift_table"] = torch.randn(1, 2, dim, generator=g) / dim ** 0.5
    lin("proj_out", out_channels * pp, dim)
    return {k: v.to(dtype) for k, v in w.items()}
