"""Flux2 DiT (FLUX.2-dev / Klein-4B / Klein-9B) forward on the B200 kernels -- BASELINE.json configs[0] family.

Host-side mirror of the reference's ``Flux2Transformer2DModel`` (apps/api/src/transformer/flux2/base/model.py:728, forward
:886-1032; dual-stream block :521-627; parallel single-stream block :449-518 with its fused processor :300-356): same
constructor config names, same state-dict keys (no biases anywhere), same ``forward(hidden_states, encoder_hidden_states,
timestep, img_ids, txt_ids, guidance, return_dict=False) -> (Tensor,)`` contract.

* dual-stream blocks: ``mmdit.dual_stream_block`` (text rows first, torch.nn.RMSNorm per head, RoPE on BOTH streams from the
  4-axis ids) with the SwiGLU feed-forward (``linear_in`` -> ``b200_swiglu`` -> ``linear_out`` with the gated residual as GEMM
  epilogue).  The three modulation modules are SHARED by all blocks (:823-834), so they are three M=1 GEMMs per forward;
* single-stream ("parallel") blocks, 5 launches each on the joint [text, image] stream: adaLN modulate -> ONE GEMM for
  q|k|v|mlp_in (``to_qkv_mlp_proj``) -> per-head RMS-norm + RoPE in place on the q, k column blocks -> SwiGLU of the mlp
  columns straight into the right part of the [S, dim + mlp] concat buffer and attention straight into its left part
  (the reference's split / chunk / cat disappear) -> ``to_out`` GEMM (K = dim + mlp) with ``h += gate * y`` as epilogue.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .. import ops
from ..flux.model import flux_rope_table
from ..lora import LoraHostMixin
from ..mmdit import JointWorkspace, StreamParams, dual_stream_block, fuse_linears, sinusoid_256


@dataclass
class Flux2Config:
    """Constructor arguments of the reference class (model.py:791-806); defaults = FLUX.2-dev."""
    patch_size: int = 1
    in_channels: int = 128
    out_channels: Optional[int] = None
    num_layers: int = 8
    num_single_layers: int = 48
    attention_head_dim: int = 128
    num_attention_heads: int = 48
    joint_attention_dim: int = 15360
    timestep_guidance_channels: int = 256
    mlp_ratio: float = 3.0
    axes_dims_rope: Tuple[int, ...] = (32, 32, 32, 32)
    rope_theta: int = 2000
    eps: float = 1e-6
    guidance_embeds: bool = True

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim


class Flux2Transformer2DModel(LoraHostMixin):
    """B200 implementation; see module docstring."""

    def __init__(self, config: Optional[Flux2Config] = None, **kwargs):
        self.config = config or Flux2Config(**kwargs)
        c = self.config
        if c.attention_head_dim != 128:
            raise ValueError("the b200 attention kernel supports attention_head_dim == 128 only")
        if sum(c.axes_dims_rope) != c.attention_head_dim:
            raise ValueError(f"axes_dims_rope {c.axes_dims_rope} must sum to attention_head_dim")
        if c.patch_size != 1 or c.timestep_guidance_channels != 256:
            raise ValueError("patch_size != 1 / timestep_guidance_channels != 256 are not implemented")
        self.mlp = int(c.inner_dim * c.mlp_ratio)
        self.w: Dict[str, torch.Tensor] = {}
        self._rope_cache: Dict[Tuple, torch.Tensor] = {}
        self._ws = None
        self._n_out = 0
        self.dtype = torch.bfloat16
        self.device = None

    @classmethod
    def from_config(cls, config, **kwargs):
        if isinstance(config, Flux2Config):
            return cls(config)
        names = Flux2Config.__dataclass_fields__.keys()
        return cls(Flux2Config(**{k: (tuple(v) if k == "axes_dims_rope" else v) for k, v in dict(config).items() if k in names}),
                   **kwargs)

    # ------------------------------------------------------------------------------------ weights
    def _linear_dims(self) -> Dict[str, Tuple[int, int]]:
        c, d, mlp = self.config, self.config.inner_dim, self.mlp
        dims = {"time_guidance_embed.timestep_embedder.linear_1": (d, 256), "time_guidance_embed.timestep_embedder.linear_2": (d, d),
                "double_stream_modulation_img.linear": (6 * d, d), "double_stream_modulation_txt.linear": (6 * d, d),
                "single_stream_modulation.linear": (3 * d, d), "x_embedder": (d, c.in_channels),
                "context_embedder": (d, c.joint_attention_dim), "norm_out.linear": (2 * d, d),
                "proj_out": (c.out_channels or c.in_channels, d)}
        if c.guidance_embeds:
            dims.update({"time_guidance_embed.guidance_embedder.linear_1": (d, 256),
                         "time_guidance_embed.guidance_embedder.linear_2": (d, d)})
        for i in range(c.num_layers):
            p = f"transformer_blocks.{i}"
            dims.update({f"{p}.attn.{n}": (d, d) for n in ("to_q", "to_k", "to_v", "to_out.0", "add_q_proj", "add_k_proj",
                                                           "add_v_proj", "to_add_out")})
            for f in ("ff", "ff_context"):
                dims[f"{p}.{f}.linear_in"], dims[f"{p}.{f}.linear_out"] = (2 * mlp, d), (d, mlp)
        for i in range(c.num_single_layers):
            p = f"single_transformer_blocks.{i}.attn"
            dims[p + ".to_qkv_mlp_proj"], dims[p + ".to_out"] = (3 * d + 2 * mlp, d), (d, d + mlp)
        return dims

    def state_dict_keys(self) -> List[str]:
        c = self.config
        keys = [m + ".weight" for m in self._linear_dims()]
        for i in range(c.num_layers):
            keys += [f"transformer_blocks.{i}.attn.{n}.weight" for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k")]
        for i in range(c.num_single_layers):
            keys += [f"single_transformer_blocks.{i}.attn.norm_q.weight", f"single_transformer_blocks.{i}.attn.norm_k.weight"]
        return keys

    def _finish_weights(self, w: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        for i in range(self.config.num_layers):
            a = f"transformer_blocks.{i}.attn"
            fuse_linears(w, a + ".to_qkv", [a + ".to_q", a + ".to_k", a + ".to_v"])
            fuse_linears(w, a + ".add_qkv", [a + ".add_q_proj", a + ".add_k_proj", a + ".add_v_proj"])
        self._n_out = w["proj_out.weight"].shape[0]
        n_pad = (-self._n_out) % 8
        if n_pad:
            w["proj_out.weight"] = F.pad(w["proj_out.weight"], (0, 0, 0, n_pad)).contiguous()
        return w

    def load_state_dict(self, state: Dict[str, torch.Tensor], device="cuda", strict: bool = True):
        want = set(self.state_dict_keys())
        missing, unexpected = sorted(want - set(state)), sorted(set(state) - want)
        if strict and (missing or unexpected):
            raise KeyError(f"state dict mismatch: missing {missing[:5]}... unexpected {unexpected[:5]}...")
        dev = torch.device(device)
        self.device = dev
        w = {k: v.detach().to(device=dev, dtype=torch.bfloat16).contiguous() for k, v in state.items() if k in want}
        self.w = self._finish_weights(w)
        return missing, unexpected

    def init_random_weights(self, device="cuda", seed: int = 1234, std: float = 0.02):
        dev = torch.device(device)
        self.device = dev
        g = torch.Generator(device=dev).manual_seed(seed)
        rnd = lambda *shape, scale=std, base=0.0: (torch.randn(*shape, generator=g, device=dev, dtype=torch.float32) * scale + base).to(torch.bfloat16)
        w = {name + ".weight": rnd(o, i) for name, (o, i) in self._linear_dims().items()}
        for k in self.state_dict_keys():
            if k not in w:
                w[k] = rnd(128, base=1.0)
        self.w = self._finish_weights(w)
        return self

    def lora_target(self, module: str):
        d = self.config.inner_dim
        head, _, leaf = module.rpartition(".")
        fused = None
        if head.endswith(".attn") and leaf in ("to_q", "to_k", "to_v") and head.startswith("transformer_blocks."):
            fused = (head + ".to_qkv", ("to_q", "to_k", "to_v").index(leaf) * d, d)
        elif head.endswith(".attn") and leaf in ("add_q_proj", "add_k_proj", "add_v_proj"):
            fused = (head + ".add_qkv", ("add_q_proj", "add_k_proj", "add_v_proj").index(leaf) * d, d)
        if fused is not None and fused[0] + ".weight" in self.w:
            return fused[0] + ".weight", fused[1], fused[2], fused[0] + ".bias"
        if module == "proj_out":
            return "proj_out.weight", 0, self._n_out, "proj_out.bias"
        if module + ".weight" in self.w and self.w[module + ".weight"].dim() == 2:
            return module + ".weight", 0, self.w[module + ".weight"].shape[0], module + ".bias"
        raise ValueError(f"Target module {module} not found in the model (or not a linear layer the b200 path adapts)")

    def parameter_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.w.values())

    # ------------------------------------------------------------------------------------ forward
    def _rope(self, txt_ids: torch.Tensor, img_ids: torch.Tensor) -> torch.Tensor:
        ids = torch.cat((txt_ids.detach().cpu().float(), img_ids.detach().cpu().float()), dim=0)
        key = (tuple(ids.shape), hash(ids.numpy().tobytes()), str(self.device))
        if key not in self._rope_cache:
            if len(self._rope_cache) > 8:
                self._rope_cache.clear()
            self._rope_cache[key] = flux_rope_table(ids, self.config.axes_dims_rope, self.device, float(self.config.rope_theta))
        return self._rope_cache[key]

    def _embed(self, t: torch.Tensor, name: str) -> torch.Tensor:
        w, p = self.w, "time_guidance_embed." + name
        h1 = ops.linear(sinusoid_256(t, self.device).to(torch.bfloat16), w[p + ".linear_1.weight"], None, epilogue=ops.EPI_SILU)
        return ops.linear(h1, w[p + ".linear_2.weight"], None)

    def single_block(self, i: int, ws, mod: Tuple[torch.Tensor, ...], rope: torch.Tensor) -> None:
        """Flux2SingleTransformerBlock.forward (model.py:479-518) in place on ws.h."""
        c, w, p = self.config, self.w, f"single_transformer_blocks.{i}.attn"
        d, H, mlp, S = c.inner_dim, c.num_attention_heads, self.mlp, ws.tokens
        shift, scale, gate = mod
        ops.adaln_zero_modulate(ws.h, scale, shift, eps=c.eps, out=ws.norm)
        ops.linear(ws.norm, w[p + ".to_qkv_mlp_proj.weight"], None, out=ws.big)                   # [S, 3d + 2 mlp]
        ops.headnorm_rope_(ws.big[:, :d], ws.big[:, d:2 * d], w[p + ".norm_q.weight"], w[p + ".norm_k.weight"], rope, H, c.eps,
                           ops.NORM_TORCH_RMS)
        ops.swiglu(ws.big[:, 3 * d:], out=ws.cat[:, d:])
        as4 = lambda t: t.view(1, S, H, 128).transpose(1, 2)
        ops.attention(as4(ws.big[:, :d]), as4(ws.big[:, d:2 * d]), as4(ws.big[:, 2 * d:3 * d]), out=as4(ws.cat[:, :d]))
        ops.linear(ws.cat, w[p + ".to_out.weight"], None, epilogue=ops.EPI_GATE_RES, out=ws.h, gate=gate)

    @torch.inference_mode()
    def forward(self, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor = None, timestep: torch.Tensor = None,
                img_ids: torch.Tensor = None, txt_ids: torch.Tensor = None, guidance: Optional[torch.Tensor] = None,
                joint_attention_kwargs=None, return_dict: bool = False, **unused):
        if not self.w:
            raise RuntimeError("weights not loaded: call load_state_dict() or init_random_weights()")
        c, w, bf, dev = self.config, self.w, torch.bfloat16, self.device
        x_in = hidden_states.to(device=dev, dtype=bf)
        enc = encoder_hidden_states.to(device=dev, dtype=bf)
        if img_ids.ndim == 3:
            img_ids = img_ids[0]
        if txt_ids.ndim == 3:
            txt_ids = txt_ids[0]
        b, n_img, _ = x_in.shape
        n_txt = enc.shape[1]
        d, H, mlp = c.inner_dim, c.num_attention_heads, self.mlp
        t = timestep.to(device=dev, dtype=bf) * 1000                     # model.py:931
        temb = self._embed(t, "timestep_embedder")
        if guidance is not None and c.guidance_embeds:
            temb = temb + self._embed(guidance.to(device=dev, dtype=bf) * 1000, "guidance_embedder")
        act = F.silu(temb)
        mod_img = ops.linear(act, w["double_stream_modulation_img.linear.weight"], None)     # [B, 6d]
        mod_txt = ops.linear(act, w["double_stream_modulation_txt.linear.weight"], None)
        mod_single = ops.linear(act, w["single_stream_modulation.linear.weight"], None)      # [B, 3d]
        mod_out = ops.linear(act, w["norm_out.linear.weight"], None)                         # [B, 2d]: scale, shift
        rope = self._rope(txt_ids, img_ids)
        S = n_txt + n_img
        ws = self._ws
        if ws is None or ws.tokens != S:
            self._ws = ws = JointWorkspace(S, d, 2 * mlp, dev, ffn_act_dim=mlp)
            ws.big = torch.empty(S, 3 * d + 2 * mlp, dtype=bf, device=dev)
            ws.cat = torch.empty(S, d + mlp, dtype=bf, device=dev)
        outs = []
        for bi in range(b):
            ops.linear(enc[bi], w["context_embedder.weight"], None, out=ws.h[:n_txt])
            ops.linear(x_in[bi], w["x_embedder.weight"], None, out=ws.h[n_txt:])
            mi, mt = mod_img[bi].chunk(6), mod_txt[bi].chunk(6)
            for i in range(c.num_layers):
                p = f"transformer_blocks.{i}"
                streams = (
                    StreamParams(slice(n_txt, None), mi, p + ".attn.to_qkv", p + ".attn.norm_q.weight", p + ".attn.norm_k.weight",
                                 p + ".attn.to_out.0", p + ".ff", rope[n_txt:], swiglu=True),
                    StreamParams(slice(0, n_txt), mt, p + ".attn.add_qkv", p + ".attn.norm_added_q.weight",
                                 p + ".attn.norm_added_k.weight", p + ".attn.to_add_out", p + ".ff_context", rope[:n_txt], swiglu=True),
                )
                dual_stream_block(w, ws, streams, H, ops.NORM_TORCH_RMS, eps=c.eps)
            ms = mod_single[bi].chunk(3)
            for i in range(c.num_single_layers):
                self.single_block(i, ws, ms, rope)
            scale, shift = mod_out[bi].chunk(2)
            ops.adaln_zero_modulate(ws.h[n_txt:], scale, shift, eps=c.eps, out=ws.norm[n_txt:])
            outs.append(ops.linear(ws.norm[n_txt:], w["proj_out.weight"], None)[:, :self._n_out])
        out = torch.stack(outs, dim=0)
        if return_dict:
            return {"sample": out}
        return (out,)

    __call__ = forward

    def eval(self):
        return self

    def to(self, *args, **kwargs):
        return self
