"""CPU ORACLE (test infrastructure -- never imported by the product package) for the Flux DiT path
(BASELINE.json configs[1]: Flux-Dev t2i 1024x1024; SURVEY.md section 8 f1 "dual-stream block variants").

A functional restatement, in plain torch on the CPU, of ``FluxTransformer2DModel.forward`` (all paths relative to
/root/reference/apps/api/src/):

    transformer/flux/base/model.py:468-657   forward                      -> :func:`flux_forward`
    transformer/flux/base/model.py:231-328   FluxTransformerBlock         -> :func:`dual_block`
    transformer/flux/base/model.py:166-228   FluxSingleTransformerBlock   -> :func:`single_block`
    transformer/flux/base/attention.py:47-116 FluxAttnProcessor           -> :func:`joint_attention`
    transformer/flux/base/model.py:331-361   FluxPosEmbed                 -> :func:`rope_table`
    engine/flux/shared.py:504-620            base_denoise loop            -> :func:`denoise`
    engine/flux/shared.py:30-55,198-215      latent packing / ids         -> :func:`pack_latents`, :func:`latent_image_ids`
    engine/flux/shared.py:58-70              calculate_shift (mu)         -> :func:`calculate_shift`

Arithmetic that lives in the un-vendored, un-pinned dependency ``diffusers`` is restated from its published
semantics (stand-in: oracle/ref_import/diffusers): AdaLayerNormZero / AdaLayerNormZeroSingle /
AdaLayerNormContinuous, CombinedTimestep(Guidance)TextProjEmbeddings, get_1d_rotary_pos_embed, apply_rotary_emb,
FeedForward("gelu-approximate") and FlowMatchEulerDiscreteScheduler (:class:`FlowMatchEuler`; the scheduler has no
in-tree twin, so its parity is UNPINNED -- the transformer forward is pinned, see below).

Run with every tensor in float32 this is the exact-math oracle; with bf16 weights / activations every torch op
rounds where the reference's module graph rounds, so on the CPU it reproduces the reference's bf16 forward bit for bit.

PINNING: oracle/make_golden.py runs the reference's own FluxTransformer2DModel (imported unmodified from
/root/reference through the stand-in) on seeded inputs -> tests/golden/flux_*.npz; tests/test_oracle_flux.py checks this
restatement against those vectors (fp32 and bf16).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Weights = Dict[str, torch.Tensor]


def linear(x: torch.Tensor, w: Weights, prefix: str) -> torch.Tensor:
    return F.linear(x, w[prefix + ".weight"], w.get(prefix + ".bias"))


# --------------------------------------------------------------------------------------------------
# embedders
# --------------------------------------------------------------------------------------------------
def sinusoid(t: torch.Tensor, dim: int = 256) -> torch.Tensor:
    """diffusers Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0): fp32 [cos | sin]."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    arg = t[:, None].float() * freqs[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


def time_text_embed(timestep: torch.Tensor, guidance: Optional[torch.Tensor], pooled: torch.Tensor, w: Weights) -> torch.Tensor:
    """CombinedTimestep[Guidance]TextProjEmbeddings (model.py:432-440 build, :537-541 call)."""
    p = "time_text_embed."
    emb = linear(F.silu(linear(sinusoid(timestep).to(pooled.dtype), w, p + "timestep_embedder.linear_1")), w,
                 p + "timestep_embedder.linear_2")
    if guidance is not None:
        g = linear(F.silu(linear(sinusoid(guidance).to(pooled.dtype), w, p + "guidance_embedder.linear_1")), w,
                   p + "guidance_embedder.linear_2")
        emb = emb + g
    txt = linear(F.silu(linear(pooled, w, p + "text_embedder.linear_1")), w, p + "text_embedder.linear_2")
    return emb + txt


def rope_table(ids: torch.Tensor, axes_dim: Sequence[int] = (16, 56, 56), theta: float = 10000.0):
    """FluxPosEmbed.forward (model.py:338-361): per axis cos/sin of pos x theta^(-2i/d) in float64, each value repeated
    for its (even, odd) channel pair, cast to float32.  Returns (cos, sin) [S, sum(axes_dim)]."""
    pos = ids.float()
    cos_out, sin_out = [], []
    for i, d in enumerate(axes_dim):
        freqs = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float64)[: d // 2] / d))
        ang = torch.outer(pos[:, i], freqs)
        cos_out.append(ang.cos().repeat_interleave(2, dim=1).float())
        sin_out.append(ang.sin().repeat_interleave(2, dim=1).float())
    return torch.cat(cos_out, dim=-1), torch.cat(sin_out, dim=-1)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """diffusers apply_rotary_emb(x [B,S,H,D], (cos, sin), sequence_dim=1): fp32 math, one cast back."""
    c, s = cos[None, :, None, :], sin[None, :, None, :]
    re, im = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-im, re], dim=-1).flatten(3)
    return (x.float() * c + rot.float() * s).to(x.dtype)


# --------------------------------------------------------------------------------------------------
# blocks
# --------------------------------------------------------------------------------------------------
def ada_modulate(x: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """``norm(x) * (1 + scale[:, None]) + shift[:, None]`` with nn.LayerNorm(elementwise_affine=False)."""
    return F.layer_norm(x, (x.shape[-1],), None, None, eps) * (1 + scale[:, None]) + shift[:, None]


def head_rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """torch.nn.RMSNorm(head_dim) on [B,S,H,D] (model.py:102-107)."""
    return F.rms_norm(x, (x.shape[-1],), weight, eps)


def sdpa(q, k, v):
    """attention/functions.py:338-377 (`sdpa` backend): q,k,v [B,H,S,D]."""
    return F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)


def joint_attention(w: Weights, p: str, heads: int, x: torch.Tensor, ctx: Optional[torch.Tensor], rope) -> Tuple:
    """FluxAttnProcessor.__call__ (attention.py:56-116): image (and text) q/k/v, per-head RMS norm, [text, image]
    concatenation, RoPE over the joint sequence, attention, split, output projections."""
    q = linear(x, w, p + ".to_q").unflatten(-1, (heads, -1))
    k = linear(x, w, p + ".to_k").unflatten(-1, (heads, -1))
    v = linear(x, w, p + ".to_v").unflatten(-1, (heads, -1))
    q = head_rms_norm(q, w[p + ".norm_q.weight"])
    k = head_rms_norm(k, w[p + ".norm_k.weight"])
    if ctx is not None:
        eq = linear(ctx, w, p + ".add_q_proj").unflatten(-1, (heads, -1))
        ek = linear(ctx, w, p + ".add_k_proj").unflatten(-1, (heads, -1))
        ev = linear(ctx, w, p + ".add_v_proj").unflatten(-1, (heads, -1))
        eq = head_rms_norm(eq, w[p + ".norm_added_q.weight"])
        ek = head_rms_norm(ek, w[p + ".norm_added_k.weight"])
        q, k, v = torch.cat([eq, q], dim=1), torch.cat([ek, k], dim=1), torch.cat([ev, v], dim=1)
    if rope is not None:
        q, k = apply_rope(q, *rope), apply_rope(k, *rope)
    o = sdpa(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3))
    o = o.permute(0, 2, 1, 3).flatten(2, 3).to(q.dtype)
    if ctx is None:
        return o
    n_txt = ctx.shape[1]
    return linear(o[:, n_txt:], w, p + ".to_out.0"), linear(o[:, :n_txt], w, p + ".to_add_out")


def feed_forward(x: torch.Tensor, w: Weights, p: str) -> torch.Tensor:
    """diffusers FeedForward(activation_fn="gelu-approximate")."""
    return linear(F.gelu(linear(x, w, p + ".net.0.proj"), approximate="tanh"), w, p + ".net.2")


def dual_block(i: int, w: Weights, heads: int, x: torch.Tensor, ctx: torch.Tensor, temb: torch.Tensor, rope):
    """FluxTransformerBlock.forward (model.py:257-328) -> (ctx, x)."""
    p = f"transformer_blocks.{i}"
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = linear(F.silu(temb), w, p + ".norm1.linear").chunk(6, dim=1)
    c_shift_msa, c_scale_msa, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = linear(
        F.silu(temb), w, p + ".norm1_context.linear").chunk(6, dim=1)
    nx = ada_modulate(x, scale_msa, shift_msa)
    nc = ada_modulate(ctx, c_scale_msa, c_shift_msa)
    ax, ac = joint_attention(w, p + ".attn", heads, nx, nc, rope)
    x = x + gate_msa.unsqueeze(1) * ax
    nx = F.layer_norm(x, (x.shape[-1],), None, None, 1e-6) * (1 + scale_mlp[:, None]) + shift_mlp[:, None]
    x = x + gate_mlp.unsqueeze(1) * feed_forward(nx, w, p + ".ff")
    ctx = ctx + c_gate_msa.unsqueeze(1) * ac
    nc = F.layer_norm(ctx, (ctx.shape[-1],), None, None, 1e-6) * (1 + c_scale_mlp[:, None]) + c_shift_mlp[:, None]
    ctx = ctx + c_gate_mlp.unsqueeze(1) * feed_forward(nc, w, p + ".ff_context")
    return ctx, x


def single_block(i: int, w: Weights, heads: int, x: torch.Tensor, ctx: torch.Tensor, temb: torch.Tensor, rope):
    """FluxSingleTransformerBlock.forward (model.py:194-228) -> (ctx, x)."""
    p = f"single_transformer_blocks.{i}"
    n_txt = ctx.shape[1]
    h = torch.cat([ctx, x], dim=1)
    shift, scale, gate = linear(F.silu(temb), w, p + ".norm.linear").chunk(3, dim=1)
    nh = ada_modulate(h, scale, shift)
    mlp = F.gelu(linear(nh, w, p + ".proj_mlp"), approximate="tanh")
    attn = joint_attention(w, p + ".attn", heads, nh, None, rope)
    h = h + gate.unsqueeze(1) * linear(torch.cat([attn, mlp], dim=2), w, p + ".proj_out")
    return h[:, :n_txt], h[:, n_txt:]


def flux_forward(hidden: torch.Tensor, enc: torch.Tensor, pooled: torch.Tensor, timestep: torch.Tensor,
                 img_ids: torch.Tensor, txt_ids: torch.Tensor, guidance: Optional[torch.Tensor], w: Weights, *,
                 heads: int, num_layers: int, num_single_layers: int, axes_dims_rope=(16, 56, 56)) -> torch.Tensor:
    """FluxTransformer2DModel.forward (model.py:468-657): hidden [B,S_img,C_in], enc [B,S_txt,joint], pooled [B,P],
    timestep [B] (already divided by 1000 by the caller, engine/flux/shared.py:548), guidance [B] or None."""
    x = linear(hidden, w, "x_embedder")
    t = timestep.to(x.dtype) * 1000
    g = guidance.to(x.dtype) * 1000 if guidance is not None else None
    temb = time_text_embed(t, g, pooled, w)
    ctx = linear(enc, w, "context_embedder")
    rope = rope_table(torch.cat((txt_ids, img_ids), dim=0), axes_dims_rope)
    for i in range(num_layers):
        ctx, x = dual_block(i, w, heads, x, ctx, temb, rope)
    for i in range(num_single_layers):
        ctx, x = single_block(i, w, heads, x, ctx, temb, rope)
    # AdaLayerNormContinuous (model.py:451-453,649): scale first, then shift
    scale, shift = linear(F.silu(temb).to(x.dtype), w, "norm_out.linear").chunk(2, dim=1)
    x = F.layer_norm(x, (x.shape[-1],), None, None, 1e-6) * (1 + scale)[:, None, :] + shift[:, None, :]
    return linear(x, w, "proj_out")


# --------------------------------------------------------------------------------------------------
# engine glue
# --------------------------------------------------------------------------------------------------
def pack_latents(latents: torch.Tensor) -> torch.Tensor:
    """engine/flux/shared.py:30-39: [B,C,H,W] -> [B,(H/2)(W/2),4C]."""
    b, c, h, w_ = latents.shape
    return latents.view(b, c, h // 2, 2, w_ // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(b, (h // 2) * (w_ // 2), c * 4)


def unpack_latents(latents: torch.Tensor, height: int, width: int, vae_scale_factor: int = 8) -> torch.Tensor:
    """engine/flux/shared.py:42-55."""
    b, _, ch = latents.shape
    h, w_ = 2 * (int(height) // (vae_scale_factor * 2)), 2 * (int(width) // (vae_scale_factor * 2))
    return latents.view(b, h // 2, w_ // 2, ch // 4, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(b, ch // 4, h, w_)


def latent_image_ids(h: int, w_: int) -> torch.Tensor:
    """engine/flux/shared.py:198-215: [h*w, 3] = (0, row, col)."""
    ids = torch.zeros(h, w_, 3)
    ids[..., 1] = ids[..., 1] + torch.arange(h)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(w_)[None, :]
    return ids.reshape(h * w_, 3)


def calculate_shift(image_seq_len, base_seq_len=256, max_seq_len=4096, base_shift=0.5, max_shift=1.15) -> float:
    """engine/flux/shared.py:58-70."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


class FlowMatchEuler:
    """diffusers FlowMatchEulerDiscreteScheduler with ``use_dynamic_shifting=True`` (FLUX.1-dev scheduler config),
    restated from the published algorithm -- PARITY UNPINNED (the dependency is not in /root/reference and the
    reference has no test or golden vector for it).  set_timesteps(sigmas=linspace(1, 1/N, N), mu) as called by
    engine/flux/t2i.py:110-135; step = Euler on the flow ODE in fp32."""

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 3.0):
        self.num_train_timesteps, self.shift = num_train_timesteps, shift
        self.order = 1

    def set_timesteps(self, num_inference_steps: int, mu: float, sigmas=None):
        if sigmas is None:
            sigmas = np.linspace(1.0, 1.0 / num_inference_steps, num_inference_steps)
        sigmas = np.array(sigmas).astype(np.float32)
        sigmas = math.exp(mu) / (math.exp(mu) + (1 / sigmas - 1) ** 1.0)  # time_shift("exponential")
        sigmas = torch.from_numpy(sigmas).to(dtype=torch.float32)
        self.timesteps = sigmas * self.num_train_timesteps
        self.sigmas = torch.cat([sigmas, torch.zeros(1)])
        self._step_index = None
        return self.timesteps

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor) -> torch.Tensor:
        if self._step_index is None:
            self._step_index = 0
        sample = sample.to(torch.float32)
        s, s_next = self.sigmas[self._step_index], self.sigmas[self._step_index + 1]
        prev = sample + (s_next - s) * model_output
        self._step_index += 1
        return prev.to(model_output.dtype)


def denoise(latents: torch.Tensor, enc, pooled, img_ids, txt_ids, guidance, w: Weights, num_inference_steps: int, **kw):
    """engine/flux/shared.py:504-620 without CFG (Flux-dev is guidance-distilled: true_cfg_scale = 1)."""
    sch = FlowMatchEuler()
    sch.set_timesteps(num_inference_steps, calculate_shift(latents.shape[1]))
    for t in sch.timesteps:
        timestep = t.expand(latents.shape[0]).to(latents.dtype)
        pred = flux_forward(latents, enc, pooled, timestep / 1000, img_ids, txt_ids, guidance, w, **kw)
        latents = sch.step(pred, t, latents)
    return latents


# --------------------------------------------------------------------------------------------------
# synthetic weights (reference state-dict keys)
# --------------------------------------------------------------------------------------------------
def make_weights(dim: int, heads: int, num_layers: int, num_single_layers: int, in_channels: int = 64,
                 joint_dim: int = 4096, pooled_dim: int = 768, guidance_embeds: bool = True, mlp_ratio: float = 4.0,
                 seed: int = 1234, dtype=torch.float32, std: float = 0.02) -> Weights:
    g = torch.Generator().manual_seed(seed)
    hd = dim // heads
    w: Weights = {}

    def lin(name, out_f, in_f, s=std):
        w[name + ".weight"] = torch.randn(out_f, in_f, generator=g) * s
        w[name + ".bias"] = torch.randn(out_f, generator=g) * std

    def gain(name):
        w[name] = 1.0 + 0.1 * torch.randn(hd, generator=g)

    emb = ["timestep_embedder"] + (["guidance_embedder"] if guidance_embeds else [])
    for e in emb:
        lin(f"time_text_embed.{e}.linear_1", dim, 256)
        lin(f"time_text_embed.{e}.linear_2", dim, dim)
    lin("time_text_embed.text_embedder.linear_1", dim, pooled_dim)
    lin("time_text_embed.text_embedder.linear_2", dim, dim)
    lin("context_embedder", dim, joint_dim)
    lin("x_embedder", dim, in_channels)
    ffn = int(dim * 4)
    for i in range(num_layers):
        p = f"transformer_blocks.{i}"
        lin(p + ".norm1.linear", 6 * dim, dim, 0.05)
        lin(p + ".norm1_context.linear", 6 * dim, dim, 0.05)
        for n in ("to_q", "to_k", "to_v", "to_out.0", "add_q_proj", "add_k_proj", "add_v_proj", "to_add_out"):
            lin(f"{p}.attn.{n}", dim, dim, 0.05)
        for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
            gain(f"{p}.attn.{n}.weight")
        for f in ("ff", "ff_context"):
            lin(f"{p}.{f}.net.0.proj", ffn, dim, 0.05)
            lin(f"{p}.{f}.net.2", dim, ffn, 0.05)
    mlp = int(dim * mlp_ratio)
    for i in range(num_single_layers):
        p = f"single_transformer_blocks.{i}"
        lin(p + ".norm.linear", 3 * dim, dim, 0.05)
        lin(p + ".proj_mlp", mlp, dim, 0.05)
        lin(p + ".proj_out", dim, dim + mlp, 0.05)
        for n in ("to_q", "to_k", "to_v"):
            lin(f"{p}.attn.{n}", dim, dim, 0.05)
        gain(f"{p}.attn.norm_q.weight")
        gain(f"{p}.attn.norm_k.weight")
    lin("norm_out.linear", 2 * dim, dim, 0.05)
    lin("proj_out", in_channels, dim)
    return {k: v.to(dtype) for k, v in w.items()}
