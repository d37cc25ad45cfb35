"""Wan DiT (2.1 / 2.2 A14B t2v) forward on the B200 kernels.

Host-side mirror of the reference's ``WanTransformer3DModel`` (apps/api/src/transformer/wan/base/model.py:1337,
forward :1684-1891, block :1101-1333, attention processor transformer/wan/base/attention.py:305-413):
same constructor config names, same diffusers-format state-dict keys, same
``forward(hidden_states, timestep, encoder_hidden_states, return_dict=False) -> (Tensor,)`` contract, so the
engine's denoise loop (engine/wan/shared/__init__.py:548-563) can call it unchanged and it can be registered as
``TRANSFORMERS_REGISTRY["wan.b200"]``.

What runs where: every FLOP of the 40 blocks, the embedders, the patch embedding and the output head goes
through libapex_b200.so (``ops``): 12 kernel launches per block --

    layernorm_modulate -> linear(QKV fused) -> rmsnorm_rope(q and k, ONE launch) -> attention
      -> linear(to_out, epilogue h += gate * y)
    layernorm(affine)  -> linear(q, epilogue * norm_q weight + row sums of squares) -> attention(logits * rstd of the row)
      -> linear(to_out, epilogue h += y)
    layernorm_modulate -> linear(ffn.0, epilogue gelu-tanh) -> linear(ffn.2, epilogue h += gate * y)

(the cross-attention's norm_q has no kernel of its own: its per-channel weight is applied by the projection's epilogue, its
row factor rsqrt(mean(q^2) + eps) by the attention kernel on the logits -- ``fuse_cross_q_norm``; 13 launches without it)

plus TWO launches per forward for the text side of every layer's cross-attention: the to_k|to_v projections of all layers
are one GEMM over the stacked weights ([L_text, 5120] x [5120, layers * 10240]) and the norm_k of all layers one batched
row kernel (they depend on the text context only; the reference recomputes them inside each block, attention.py:345-352).

torch is used for device buffers and for O(dim) glue (the [B,6,dim] modulation table add, SiLU on the
[B,dim] time embedding, patchify / unpatchify reshapes).  Activations live in preallocated workspaces that
are reused by all layers (HBM layout: token-major [S, channels] bf16; q|k|v as column blocks of one
[S, 3*dim] buffer so the attention kernel reads them through strided TMA maps without a transpose).
"""
from __future__ import annotations

import os

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from .. import ops
from ..lora import LoraHostMixin
from ..parallel import ParallelContext
from .rope import wan_rope_table_bf16


@dataclass
class WanConfig:
    """Constructor arguments of the reference class (model.py:1389-1411); defaults = Wan 2.1/2.2 A14B."""
    patch_size: Tuple[int, int, int] = (1, 2, 2)
    num_attention_heads: int = 40
    attention_head_dim: int = 128
    in_channels: int = 16
    out_channels: int = 16
    text_dim: int = 4096
    freq_dim: int = 256
    ffn_dim: int = 13824
    num_layers: int = 40
    cross_attn_norm: bool = True
    qk_norm: Optional[str] = "rms_norm_across_heads"
    eps: float = 1e-6
    rope_max_seq_len: int = 1024

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim


class _Workspace:
    """Activation buffers shared by all layers of one forward (sized for [tokens, dim])."""

    def __init__(self, tokens: int, ctx_tokens: int, cfg: WanConfig, device):
        d, bf = cfg.inner_dim, torch.bfloat16
        self.tokens, self.ctx_tokens = tokens, ctx_tokens
        self.norm = torch.empty(tokens, d, dtype=bf, device=device)
        self.qkv = torch.empty(tokens, 3 * d, dtype=bf, device=device)
        self.attn = torch.empty(tokens, d, dtype=bf, device=device)
        self.ffn = torch.empty(tokens, cfg.ffn_dim, dtype=bf, device=device)
        self.kv_ctx = torch.empty(ctx_tokens, cfg.num_layers * 2 * d, dtype=bf, device=device)   # K|V of EVERY layer's cross-attention
        # per-row, per-column-tile sums of squares of the cross-attention query (ops.linear_normw -> ops.attention(q_norm=...))
        self.q_sumsq = torch.empty(tokens * ((d + 63) // 64), dtype=torch.float32, device=device)


class WanTransformer3DModel(LoraHostMixin):
    """B200 implementation; see module docstring.  Not an nn.Module: weights are a flat dict of bf16 CUDA
    tensors keyed like the reference's state dict (``blocks.N.attn1.to_q.weight`` ...)."""

    def __init__(self, config: Optional[WanConfig] = None, **kwargs):
        self.config = config or WanConfig(**kwargs)
        if self.config.attention_head_dim != 128:
            raise ValueError("the b200 attention kernel supports attention_head_dim == 128 only")
        if self.config.qk_norm != "rms_norm_across_heads":
            raise ValueError("only qk_norm='rms_norm_across_heads' (the Wan default) is implemented")
        self.w: Dict[str, torch.Tensor] = {}
        self._rope_cache: Dict[Tuple, torch.Tensor] = {}
        self._ws: Optional[_Workspace] = None
        self.dtype = torch.bfloat16
        self.device = None
        self._use_graph = False
        self._graphs: Dict[Tuple, object] = {}
        #: norm_q of the cross-attention folded into its projection and the attention logits (B200_WAN_FUSE_Q_NORM=0: the
        #: three-kernel form projection -> RMS-norm -> attention, the A/B partner)
        self.fuse_cross_q_norm = os.environ.get("B200_WAN_FUSE_Q_NORM", "1") != "0"

    # ------------------------------------------------------------------------------------ weights
    @classmethod
    def from_config(cls, config, **kwargs) -> "WanTransformer3DModel":
        if isinstance(config, WanConfig):
            return cls(config)
        names = WanConfig.__dataclass_fields__.keys()
        return cls(WanConfig(**{k: (tuple(v) if k == "patch_size" else v) for k, v in dict(config).items()
                                if k in names}), **kwargs)

    def state_dict_keys(self):
        c = self.config
        keys = ["patch_embedding.weight", "patch_embedding.bias", "scale_shift_table", "proj_out.weight",
                "proj_out.bias"]
        for n in ("time_embedder.linear_1", "time_embedder.linear_2", "time_proj", "text_embedder.linear_1",
                  "text_embedder.linear_2"):
            keys += [f"condition_embedder.{n}.weight", f"condition_embedder.{n}.bias"]
        for i in range(c.num_layers):
            p = f"blocks.{i}"
            keys.append(p + ".scale_shift_table")
            for a in ("attn1", "attn2"):
                for n in ("to_q", "to_k", "to_v", "to_out.0"):
                    keys += [f"{p}.{a}.{n}.weight", f"{p}.{a}.{n}.bias"]
                keys += [f"{p}.{a}.norm_q.weight", f"{p}.{a}.norm_k.weight"]
            if c.cross_attn_norm:
                keys += [p + ".norm2.weight", p + ".norm2.bias"]
            for n in ("ffn.net.0.proj", "ffn.net.2"):
                keys += [f"{p}.{n}.weight", f"{p}.{n}.bias"]
        return keys

    def load_state_dict(self, state: Dict[str, torch.Tensor], device="cuda", strict: bool = True):
        """Takes the reference's (diffusers-format) state dict; casts to bf16 on ``device`` (the blanket
        ``module.to(dtype)`` of mixins/to_mixin.py:358) and fuses q|k|v (self) and k|v (cross) projections."""
        want = set(self.state_dict_keys())
        missing = sorted(want - set(state))
        unexpected = sorted(k for k in set(state) - want if "norm_added_q" not in k)
        if strict and (missing or unexpected):
            raise KeyError(f"state dict mismatch: missing {missing[:5]}... unexpected {unexpected[:5]}...")
        dev = torch.device(device)
        self.device = dev
        w = {k: v.detach().to(device=dev, dtype=torch.bfloat16).contiguous() for k, v in state.items() if k in want}
        c = self.config
        pp = c.patch_size[0] * c.patch_size[1] * c.patch_size[2]
        w["patch_embedding.weight"] = w["patch_embedding.weight"].reshape(c.inner_dim, c.in_channels * pp).contiguous()
        for i in range(c.num_layers):
            p = f"blocks.{i}"
            a1, a2 = p + ".attn1", p + ".attn2"
            w[a1 + ".to_qkv.weight"] = torch.cat([w.pop(a1 + ".to_q.weight"), w.pop(a1 + ".to_k.weight"),
                                                  w.pop(a1 + ".to_v.weight")], dim=0).contiguous()
            w[a1 + ".to_qkv.bias"] = torch.cat([w.pop(a1 + ".to_q.bias"), w.pop(a1 + ".to_k.bias"),
                                                w.pop(a1 + ".to_v.bias")], dim=0).contiguous()
            w[a2 + ".to_kv.weight"] = torch.cat([w.pop(a2 + ".to_k.weight"), w.pop(a2 + ".to_v.weight")],
                                                dim=0).contiguous()
            w[a2 + ".to_kv.bias"] = torch.cat([w.pop(a2 + ".to_k.bias"), w.pop(a2 + ".to_v.bias")], dim=0).contiguous()
        self.w = w
        self._stack_shared_layouts()
        return missing, unexpected

    def _stack_shared_layouts(self) -> None:
        """Storage layouts that let one launch serve several call sites; the per-layer names stay valid as VIEWS (LoRA merges
        into ``blocks.i.attn2.to_kv`` rows keep working in place):
          * ``blocks.i.attn1.norm_qk.weight`` [2, dim] = norm_q | norm_k  -> q and k are normalised + rotated in one launch;
          * ``attn2_kv_all.weight`` [layers * 2 dim, dim] (+ bias) and ``attn2_norm_k_all.weight`` [layers, dim] -> the text side
            of every layer's cross-attention is one GEMM + one row kernel per forward."""
        c, w = self.config, self.w
        L = c.num_layers
        for i in range(L):
            a1 = f"blocks.{i}.attn1"
            qk = torch.stack([w[a1 + ".norm_q.weight"], w[a1 + ".norm_k.weight"]], dim=0).contiguous()
            w[a1 + ".norm_qk.weight"], w[a1 + ".norm_q.weight"], w[a1 + ".norm_k.weight"] = qk, qk[0], qk[1]
        kv_w = torch.cat([w[f"blocks.{i}.attn2.to_kv.weight"] for i in range(L)], dim=0).contiguous()
        kv_b = torch.cat([w[f"blocks.{i}.attn2.to_kv.bias"] for i in range(L)], dim=0).contiguous()
        nk = torch.stack([w[f"blocks.{i}.attn2.norm_k.weight"] for i in range(L)], dim=0).contiguous()
        d2 = 2 * c.inner_dim
        for i in range(L):
            w[f"blocks.{i}.attn2.to_kv.weight"] = kv_w[i * d2:(i + 1) * d2]
            w[f"blocks.{i}.attn2.to_kv.bias"] = kv_b[i * d2:(i + 1) * d2]
            w[f"blocks.{i}.attn2.norm_k.weight"] = nk[i]
        w["attn2_kv_all.weight"], w["attn2_kv_all.bias"], w["attn2_norm_k_all.weight"] = kv_w, kv_b, nk

    def init_random_weights(self, device="cuda", seed: int = 1234, std: float = 0.02):
        """Synthetic weights of the architecture's shapes generated ON the device (bench.py; there are no
        checkpoints offline): N(0, std^2) linears, randn/sqrt(dim) modulation tables, 1 + N(0, std^2) norm gains."""
        dev = torch.device(device)
        self.device = dev
        g = torch.Generator(device=dev).manual_seed(seed)
        c, d, bf = self.config, self.config.inner_dim, torch.bfloat16
        pp = c.patch_size[0] * c.patch_size[1] * c.patch_size[2]
        w: Dict[str, torch.Tensor] = {}

        def rnd(*shape, scale=std, base=0.0):
            return (torch.randn(*shape, generator=g, device=dev, dtype=torch.float32) * scale + base).to(bf)

        def lin(name, out_f, in_f):
            w[name + ".weight"], w[name + ".bias"] = rnd(out_f, in_f), rnd(out_f)

        lin("patch_embedding", d, c.in_channels * pp)
        lin("condition_embedder.time_embedder.linear_1", d, c.freq_dim)
        lin("condition_embedder.time_embedder.linear_2", d, d)
        lin("condition_embedder.time_proj", 6 * d, d)
        lin("condition_embedder.text_embedder.linear_1", d, c.text_dim)
        lin("condition_embedder.text_embedder.linear_2", d, d)
        for i in range(c.num_layers):
            p = f"blocks.{i}"
            w[p + ".scale_shift_table"] = rnd(1, 6, d, scale=d ** -0.5)
            lin(p + ".attn1.to_qkv", 3 * d, d)
            lin(p + ".attn1.to_out.0", d, d)
            lin(p + ".attn2.to_q", d, d)
            lin(p + ".attn2.to_kv", 2 * d, d)
            lin(p + ".attn2.to_out.0", d, d)
            for a in ("attn1", "attn2"):
                w[f"{p}.{a}.norm_q.weight"], w[f"{p}.{a}.norm_k.weight"] = rnd(d, base=1.0), rnd(d, base=1.0)
            if c.cross_attn_norm:
                w[p + ".norm2.weight"], w[p + ".norm2.bias"] = rnd(d, base=1.0), rnd(d)
            lin(p + ".ffn.net.0.proj", c.ffn_dim, d)
            lin(p + ".ffn.net.2", d, c.ffn_dim)
        w["scale_shift_table"] = rnd(1, 2, d, scale=d ** -0.5)
        lin("proj_out", c.out_channels * pp, d)
        self.w = w
        self._stack_shared_layouts()
        return self

    def lora_target(self, module: str):
        """LoRA target module name (reference state-dict naming, e.g. ``blocks.3.attn1.to_k``) -> (weight key in
        ``self.w``, first row, rows, bias key): resolves the q|k|v / k|v fusion done by ``load_state_dict``."""
        d = self.config.inner_dim
        head, _, leaf = module.rpartition(".")
        fused = None
        if head.endswith(".attn1") and leaf in ("to_q", "to_k", "to_v"):
            fused = (head + ".to_qkv", ("to_q", "to_k", "to_v").index(leaf) * d)
        elif head.endswith(".attn2") and leaf in ("to_k", "to_v"):
            fused = (head + ".to_kv", ("to_k", "to_v").index(leaf) * d)
        if fused is not None and fused[0] + ".weight" in self.w:
            return fused[0] + ".weight", fused[1], d, fused[0] + ".bias"
        if module + ".weight" in self.w and self.w[module + ".weight"].dim() == 2 and module != "patch_embedding":
            return module + ".weight", 0, self.w[module + ".weight"].shape[0], module + ".bias"
        raise ValueError(f"Target module {module} not found in the model (or not a linear layer the b200 path adapts)")

    def parameter_bytes(self) -> int:
        """bytes of distinct weight storage (the stacked layouts alias their per-layer views)"""
        seen, total = set(), 0
        for t in self.w.values():
            st = t.untyped_storage()
            if st.data_ptr() not in seen:
                seen.add(st.data_ptr())
                total += st.nbytes()
        return total

    # ------------------------------------------------------------------------------------ pieces
    def _rope(self, grid: Tuple[int, int, int]) -> torch.Tensor:
        key = (grid, str(self.device))
        if key not in self._rope_cache:  # the reference caches its table too (model.py:1584-1629)
            self._rope_cache[key] = wan_rope_table_bf16(self.config.attention_head_dim, grid, self.device,
                                                        self.config.rope_max_seq_len)
        return self._rope_cache[key]

    def _workspace(self, tokens: int, ctx_tokens: int) -> _Workspace:
        ws = self._ws
        if ws is None or ws.tokens != tokens or ws.ctx_tokens != ctx_tokens:
            self._ws = ws = _Workspace(tokens, ctx_tokens, self.config, self.device)
        return ws

    def condition_embed(self, timestep: torch.Tensor, text: torch.Tensor):
        """model.py:773-823 -> temb [B,dim], timestep_proj [B,6,dim], context [B,L,dim]."""
        c, w, p = self.config, self.w, "condition_embedder."
        half = c.freq_dim // 2
        freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=self.device) / half)
        arg = timestep.to(self.device)[:, None].float() * freqs[None, :]
        ts = torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1).to(torch.bfloat16)
        t1 = F.silu(ops.linear(ts, w[p + "time_embedder.linear_1.weight"], w[p + "time_embedder.linear_1.bias"]))
        temb = ops.linear(t1, w[p + "time_embedder.linear_2.weight"], w[p + "time_embedder.linear_2.bias"])
        tproj = ops.linear(F.silu(temb), w[p + "time_proj.weight"], w[p + "time_proj.bias"])
        x1 = ops.linear(text, w[p + "text_embedder.linear_1.weight"], w[p + "text_embedder.linear_1.bias"],
                        epilogue=ops.EPI_GELU_TANH)
        ctx = ops.linear(x1, w[p + "text_embedder.linear_2.weight"], w[p + "text_embedder.linear_2.bias"])
        return temb, tproj.unflatten(1, (6, -1)), ctx

    def patchify(self, latents: torch.Tensor, lo: int = 0, hi: Optional[int] = None) -> torch.Tensor:
        """Conv3d(kernel = stride = patch) (model.py:1748-1749) as a GEMM over [tokens, C*pt*ph*pw]; only the
        token range [lo, hi) is embedded (token shards of the sequence-parallel group)."""
        c = self.config
        b, ch, f, h, w_ = latents.shape
        pt, ph, pw = c.patch_size
        x = latents.reshape(b, ch, f // pt, pt, h // ph, ph, w_ // pw, pw).permute(0, 2, 4, 6, 1, 3, 5, 7)
        x = x.reshape(b, (f // pt) * (h // ph) * (w_ // pw), ch * pt * ph * pw)[:, lo:hi].contiguous()
        return ops.linear(x, self.w["patch_embedding.weight"], self.w["patch_embedding.bias"])

    def block(self, i: int, h: torch.Tensor, ctx: torch.Tensor, temb6: torch.Tensor, rope: torch.Tensor,
              ws: _Workspace, par: Optional[ParallelContext] = None) -> None:
        """One WanTransformerBlock (model.py:1101-1333) on ONE batch element; h [S_local, dim] (this rank's
        token shard; the whole sequence when not sequence-parallel) is updated in place."""
        c, w, p = self.config, self.w, f"blocks.{i}"
        d, heads, hd, eps = c.inner_dim, c.num_attention_heads, c.attention_head_dim, c.eps
        S = h.shape[0]
        # (scale_shift_table + temb.float()).to(bf16).chunk(6)   model.py:1130-1132
        mod = (w[p + ".scale_shift_table"][0].float() + temb6.float()).to(torch.bfloat16)
        shift_msa, scale_msa, gate_msa, c_shift, c_scale, c_gate = mod.unbind(0)

        # 1. self-attention
        ops.layernorm_modulate(h, scale_msa, shift_msa, eps=eps, out=ws.norm)
        fused_exchange = par is not None and par.sp_size > 1 and par.use_p2p
        if fused_exchange:   # q/k norm + RoPE happen inside the scatter kernels below
            ops.linear(ws.norm, w[p + ".attn1.to_qkv.weight"], w[p + ".attn1.to_qkv.bias"], out=ws.qkv)
        else:                # attention.py:345-370 as one C call (projection, then norm + RoPE on q and k in place)
            ops.qkv_rmsnorm_rope(ws.norm, w[p + ".attn1.to_qkv.weight"], w[p + ".attn1.to_qkv.bias"],
                                 w[p + ".attn1.norm_q.weight"], w[p + ".attn1.norm_k.weight"], rope, heads, eps, ws.qkv)
        q, k, v = ws.qkv[:, :d], ws.qkv[:, d:2 * d], ws.qkv[:, 2 * d:]
        as4 = lambda t, n: t.view(1, n, -1, hd).transpose(1, 2)  # [1,H,n,hd] strided view
        attn_out = ws.attn
        if fused_exchange:
            # Ulysses exchange fused into the kernels over NVLink peer memory (parallel.PeerExchange):
            # norm+RoPE(q,k) and a copy of v store straight into the head-owner's planes, the attention epilogue
            # stores straight into the token-owner's [S/P, dim] buffer; two device-side barriers per layer.
            ex = par.peer_exchange(S * par.sp_size, heads, hd, h.device)
            ops.rmsnorm_rope_scatter(q, w[p + ".attn1.norm_q.weight"], rope, heads, eps, ex.qkv_peers, ex.P, 0, ex.row0)
            ops.rmsnorm_rope_scatter(k, w[p + ".attn1.norm_k.weight"], rope, heads, eps, ex.qkv_peers, ex.P, ex.plane,
                                     ex.row0)
            ops.rmsnorm_rope_scatter(v, None, None, heads, eps, ex.qkv_peers, ex.P, 2 * ex.plane, ex.row0, norm=False)
            ex.barrier(0)
            ops.attention_scatter(as4(ex.qkv[0], ex.tokens_total), as4(ex.qkv[1], ex.tokens_total),
                                  as4(ex.qkv[2], ex.tokens_total), ex.o_peers, ex.P, ex.n_local, ex.head_off, d)
            ex.barrier(1)
            attn_out = ex.o
        elif par is None or par.sp_size == 1:
            ops.attention(as4(q, S), as4(k, S), as4(v, S), out=as4(ws.attn, S))
        else:
            # Ulysses: tokens -> heads, global attention over my heads, heads -> tokens
            qkv_h = par.tokens_to_heads(ws.qkv, heads, hd)          # [3, S_total, (H/P)*hd]
            S_total = qkv_h.shape[1]
            o_h = torch.empty_like(qkv_h[0])
            ops.attention(as4(qkv_h[0], S_total), as4(qkv_h[1], S_total), as4(qkv_h[2], S_total),
                          out=as4(o_h, S_total))
            par.heads_to_tokens(o_h, out=ws.attn)
        ops.linear(attn_out, w[p + ".attn1.to_out.0.weight"], w[p + ".attn1.to_out.0.bias"],
                   epilogue=ops.EPI_GATE_RES, out=h, gate=gate_msa)

        # 2. cross-attention (K/V from the text context, no RoPE)
        if c.cross_attn_norm:
            ops.layernorm_modulate(h, ln_weight=w[p + ".norm2.weight"], ln_bias=w[p + ".norm2.bias"], eps=eps,
                                   out=ws.norm)
            xn = ws.norm
        else:
            xn = h
        fuse_q_norm = self.fuse_cross_q_norm and d >= 64
        if fuse_q_norm:
            # norm_q folded into the projection (per-channel weight in the GEMM epilogue, row statistics on the side) and into the
            # attention logits (row factor): no separate pass over q -- 12 launches per block
            q2 = ws.attn
            q_parts = ops.linear_normw(xn, w[p + ".attn2.to_q.weight"], w[p + ".attn2.to_q.bias"], w[p + ".attn2.norm_q.weight"], q2,
                                       ws.q_sumsq)
        else:
            q2 = ops.linear(xn, w[p + ".attn2.to_q.weight"], w[p + ".attn2.to_q.bias"], out=ws.attn)
            ops.rmsnorm_rope_(q2, w[p + ".attn2.norm_q.weight"], None, heads, eps)
        L = ctx.shape[0]
        # K | V of this layer's cross-attention: column block i of ws.kv_ctx, projected + normalised for ALL layers by
        # ``text_kv`` at the top of the forward
        k2, v2 = ws.kv_ctx[:, i * 2 * d:i * 2 * d + d], ws.kv_ctx[:, i * 2 * d + d:(i + 1) * 2 * d]
        o2 = ws.norm  # norm output is dead once q2 exists
        ops.attention(as4(q2, S), as4(k2, L), as4(v2, L), out=as4(o2, S),
                      q_norm=(ws.q_sumsq, q_parts, d, eps) if fuse_q_norm else None)
        ops.linear(o2, w[p + ".attn2.to_out.0.weight"], w[p + ".attn2.to_out.0.bias"], epilogue=ops.EPI_GATE_RES,
                   out=h, gate=None)

        # 3. feed-forward
        ops.layernorm_modulate(h, c_scale, c_shift, eps=eps, out=ws.norm)
        ops.mlp_gelu_(h, ws.norm, w[p + ".ffn.net.0.proj.weight"], w[p + ".ffn.net.0.proj.bias"],
                      w[p + ".ffn.net.2.weight"], w[p + ".ffn.net.2.bias"], c_gate, ws.ffn)

    def text_kv(self, ctx: torch.Tensor, ws: _Workspace) -> None:
        """to_k | to_v (attention.py:346-347) and norm_k (:349-352) of EVERY layer's cross-attention on the text context [L, dim]
        -> ws.kv_ctx [L, layers * 2 dim]: one GEMM over the stacked weights, one batched row kernel."""
        c, w = self.config, self.w
        ops.linear(ctx, w["attn2_kv_all.weight"], w["attn2_kv_all.bias"], out=ws.kv_ctx)
        ops.rmsnorm_rope_batched_(ws.kv_ctx[:, :c.inner_dim], w["attn2_norm_k_all.weight"], None, c.num_attention_heads, c.eps,
                                  c.num_layers, 2 * c.inner_dim)

    # ------------------------------------------------------------------------------------ CUDA graph
    def enable_cuda_graph(self, enabled: bool = True) -> None:
        """Replay the whole forward (12 launches x layers + glue) from ONE CUDA graph per input shape.  Every
        kernel goes through the C ABI on the current stream with host-built TMA descriptors passed by value, weights and
        workspaces are static, nothing synchronises with the host -> capture once (after two warm-up runs), replay for every
        later step, cond and uncond alike (the text embedding is a graph INPUT).  Bit-identical to eager execution
        (tests/test_gpu_parity.py::test_wan_forward_cuda_graph_replay_is_bit_identical).  Single-GPU forwards only: the
        sequence-parallel path contains collectives / symmetric-memory barriers and stays eager."""
        self._use_graph = bool(enabled)
        if not enabled:
            self._graphs.clear()

    def _graphed(self, hidden_states: torch.Tensor, timestep: torch.Tensor, text: torch.Tensor) -> torch.Tensor:
        from ..graph import GraphedCallable

        key = (tuple(hidden_states.shape), tuple(text.shape), tuple(timestep.shape), str(timestep.dtype))
        g = self._graphs.get(key)
        if g is None:
            g = GraphedCallable(lambda h, t, e: self._forward_eager(h, t, e)[0], [hidden_states, timestep, text])
            self._graphs[key] = g
        # the graph's output buffer is static: hand out a copy (the engine keeps `cond` alive across the `uncond` forward)
        return g(hidden_states, timestep, text).clone()

    # ------------------------------------------------------------------------------------ forward
    @torch.inference_mode()
    def forward(self, hidden_states: torch.Tensor, timestep: torch.Tensor, encoder_hidden_states: torch.Tensor,
                return_dict: bool = False, **unused):
        """hidden_states [B,C,F,H,W] (bf16), timestep [B] int64, encoder_hidden_states [B,L,text_dim] ->
        ([B,C_out,F,H,W],); see ``_forward_eager``.  With ``enable_cuda_graph()`` single-GPU forwards replay a captured graph."""
        par = unused.get("parallel", None)
        if self._use_graph and (par is None or par.world_size == 1):
            if not self.w:
                raise RuntimeError("weights not loaded: call load_state_dict() or init_random_weights()")
            out = self._graphed(hidden_states.to(device=self.device, dtype=torch.bfloat16).contiguous(),
                                timestep.to(self.device).contiguous(),
                                encoder_hidden_states.to(device=self.device, dtype=torch.bfloat16).contiguous())
            return {"sample": out} if return_dict else (out,)
        return self._forward_eager(hidden_states, timestep, encoder_hidden_states, return_dict=return_dict, **unused)

    @torch.inference_mode()
    def _forward_eager(self, hidden_states: torch.Tensor, timestep: torch.Tensor, encoder_hidden_states: torch.Tensor,
                       return_dict: bool = False, **unused):
        """hidden_states [B,C,F,H,W] (bf16), timestep [B] int64, encoder_hidden_states [B,L,text_dim] ->
        ([B,C_out,F,H,W],).  Batch elements run one after another, exactly like the reference runs cond and
        uncond as two B=1 forwards (engine/wan/shared/__init__.py:548-563)."""
        if not self.w:
            raise RuntimeError("weights not loaded: call load_state_dict() or init_random_weights()")
        c = self.config
        hidden_states = hidden_states.to(device=self.device, dtype=torch.bfloat16)
        encoder_hidden_states = encoder_hidden_states.to(device=self.device, dtype=torch.bfloat16)
        b, ch, f, hh, ww = hidden_states.shape
        pt, ph, pw = c.patch_size
        if f % pt or hh % ph or ww % pw:
            raise ValueError(f"Input dims must be divisible by patch_size. Got (T,H,W)=({f},{hh},{ww}), patch={c.patch_size}")
        grid = (f // pt, hh // ph, ww // pw)
        S = grid[0] * grid[1] * grid[2]
        par: ParallelContext = unused.pop("parallel", None) or ParallelContext.single()
        lo, hi = par.shard_bounds(S)
        rope = self._rope(grid)[lo:hi]
        tokens = self.patchify(hidden_states, lo, hi)               # [B,S_local,dim]
        temb, temb6, ctx = self.condition_embed(timestep, encoder_hidden_states)
        ws = self._workspace(hi - lo, ctx.shape[1])
        outs = []
        for bi in range(b):
            h = tokens[bi]
            self.text_kv(ctx[bi], ws)
            for i in range(c.num_layers):
                self.block(i, h, ctx[bi], temb6[bi], rope, ws, par)
            # output head (model.py:1841-1868): (table + temb) in bf16, modulated norm, proj_out
            shift, scale = (self.w["scale_shift_table"][0] + temb[bi][None, :]).unbind(0)
            ops.layernorm_modulate(h, scale.contiguous(), shift.contiguous(), eps=c.eps, out=ws.norm)
            y_local = ops.linear(ws.norm, self.w["proj_out.weight"], self.w["proj_out.bias"])
            outs.append(par.gather_tokens(y_local))
        y = torch.stack(outs, dim=0)                                # [B,S,out*pp]
        y = y.reshape(b, grid[0], grid[1], grid[2], pt, ph, pw, -1).permute(0, 7, 1, 4, 2, 5, 3, 6)
        out = y.flatten(6, 7).flatten(4, 5).flatten(2, 3)
        if return_dict:
            return {"sample": out}
        return (out,)

    __call__ = forward

    # reference API no-ops kept so engine code calling them does not break (model.py:1524, 1645)
    def set_chunking_profile(self, profile_name: str) -> None:
        """Chunking profiles exist to fit small GPUs (model.py:1485-1554); on 180 GB they are pure overhead."""
        if profile_name not in ("none", "light", "balanced", "aggressive"):
            raise ValueError(f"Unknown chunking profile: {profile_name}")

    def eval(self):
        return self

    def to(self, *args, **kwargs):
        return self
