"""HunyuanVideo-1.5 DiT forward on the B200 kernels -- BASELINE.json configs[4] (I2V / T2V 720p x 129 frames).

Host-side mirror of the reference's ``HunyuanVideo15Transformer3DModel``
(apps/api/src/transformer/hunyuanvideo15/base/model.py:696, forward :964-1165; dual-stream block :617-694; attention
processor :83-171; token refiner :297-496): same constructor config names, same diffusers-format state-dict keys, same
``forward(hidden_states, timestep, encoder_hidden_states, encoder_attention_mask, encoder_hidden_states_2=...,
encoder_attention_mask_2=..., image_embeds=..., return_dict=False) -> (Tensor,)`` contract as the engine's loop calls it
(engine/hunyuanvideo15/t2v.py:234-325).

What runs where
  * 54 dual-stream blocks (99.9 % of the FLOPs; 91 % of them the joint attention over ~120.8k tokens): ``mmdit.dual_stream_block``
    -- one residual stream [S_latent + S_text, 2048] with the latent rows first (the reference concatenates [latent, encoder],
    :146-148), per-head InplaceRMSNorm + RoPE on the latent rows only (``ops.NORM_INPLACE_RMS``; table rounded to bf16 first as
    efficiency/ops.py:201-202 does), all 54 x 2 AdaLayerNormZero linears + the head's AdaLayerNormContinuous computed by ONE
    weight-streaming GEMM per forward.
  * condition embedders (text token refiner with its two attention blocks, ByT5 projection, image projection): same kernels.
    The refiner's key-padding mask (:386-402) is honoured by COMPACTION: the valid tokens are gathered first and the refiner
    runs on them without a mask -- identical for the valid rows, and the reference replaces the padded rows by zeros right
    after (:1068-1075), so nothing else is observable.  Token-wise embedders also run on valid tokens only.
  * torch: device buffers, gathers / concatenation of the <= 2k condition tokens, the masked mean for the pooled text, the
    [n, dim] ``cond_type_embed`` adds, SiLU on [B, dim] embeddings, patchify / unpatchify reshapes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .. import ops
from ..lora import LoraHostMixin
from ..mmdit import JointWorkspace, StreamParams, dual_stream_block, fuse_linears, sinusoid_256


@dataclass
class HunyuanVideo15Config:
    """Constructor arguments of the reference class (model.py:783-808); defaults = HunyuanVideo-1.5 (8.3 B)."""
    in_channels: int = 65
    out_channels: int = 32
    num_attention_heads: int = 16
    attention_head_dim: int = 128
    num_layers: int = 54
    num_refiner_layers: int = 2
    mlp_ratio: float = 4.0
    patch_size: int = 1
    patch_size_t: int = 1
    qk_norm: str = "rms_norm"
    text_embed_dim: int = 3584
    text_embed_2_dim: int = 1472
    image_embed_dim: int = 1152
    rope_theta: float = 256.0
    rope_axes_dim: Tuple[int, int, int] = (16, 56, 56)
    use_meanflow: bool = False

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim


def hy15_rope_table(grid: Tuple[int, int, int], rope_dim, theta: float, device) -> torch.Tensor:
    """HunyuanVideo15RotaryPosEmbed.forward (model.py:514-541) + the table handling of apply_cos_sin_rope_inplace
    (efficiency/ops.py:201-210): float32 frequencies (diffusers' default freqs_dtype), cos/sin cast to bf16, one entry per
    channel pair -> fp32 [F*H*W, head_dim/2, 2] holding the bf16-rounded (cos, sin)."""
    axes = [torch.arange(0, n, dtype=torch.float32) for n in grid]
    mesh = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=0)
    cs = []
    for i, d in enumerate(rope_dim):
        freqs = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float32)[: d // 2] / d))
        ang = torch.outer(mesh[i].reshape(-1), freqs)
        cs.append(torch.stack([ang.cos().float().bfloat16().float(), ang.sin().float().bfloat16().float()], dim=-1))
    return torch.cat(cs, dim=1).contiguous().to(device)


class HunyuanVideo15Transformer3DModel(LoraHostMixin):
    """B200 implementation; see module docstring."""

    def __init__(self, config: Optional[HunyuanVideo15Config] = None, **kwargs):
        self.config = config or HunyuanVideo15Config(**kwargs)
        c = self.config
        if c.attention_head_dim != 128:
            raise ValueError("the b200 attention kernel supports attention_head_dim == 128 only")
        if sum(c.rope_axes_dim) != c.attention_head_dim:
            raise ValueError(f"rope_axes_dim {c.rope_axes_dim} must sum to attention_head_dim")
        if c.qk_norm != "rms_norm":
            raise ValueError("only qk_norm='rms_norm' is implemented")
        if c.use_meanflow:
            raise ValueError("use_meanflow (super-resolution checkpoints) is not implemented on the b200 path")
        self.w: Dict[str, torch.Tensor] = {}
        self._mod_rows: Dict[str, Tuple[int, int]] = {}
        self._rope_cache: Dict[Tuple, torch.Tensor] = {}
        self._cond_plans: Dict[Tuple, dict] = {}      # per-prompt timestep-independent condition work (_condition_plan)
        self._ws: Optional[JointWorkspace] = None
        self._k_pad = 0
        self.dtype = torch.bfloat16
        self.device = None

    @classmethod
    def from_config(cls, config, **kwargs):
        if isinstance(config, HunyuanVideo15Config):
            return cls(config)
        names = HunyuanVideo15Config.__dataclass_fields__.keys()
        return cls(HunyuanVideo15Config(**{k: (tuple(v) if k == "rope_axes_dim" else v) for k, v in dict(config).items()
                                           if k in names}), **kwargs)

    # ------------------------------------------------------------------------------------ weights
    def _modulation_layout(self) -> List[Tuple[str, int]]:
        c, d = self.config, self.config.inner_dim
        lay = []
        for i in range(c.num_layers):
            lay += [(f"transformer_blocks.{i}.norm1.linear", 6 * d), (f"transformer_blocks.{i}.norm1_context.linear", 6 * d)]
        lay.append(("norm_out.linear", 2 * d))
        return lay

    def state_dict_keys(self) -> List[str]:
        c = self.config
        lin = ["image_embedder.linear_1", "image_embedder.linear_2", "context_embedder.proj_in",
               "context_embedder.time_text_embed.timestep_embedder.linear_1",
               "context_embedder.time_text_embed.timestep_embedder.linear_2",
               "context_embedder.time_text_embed.text_embedder.linear_1",
               "context_embedder.time_text_embed.text_embedder.linear_2", "context_embedder_2.linear_1",
               "context_embedder_2.linear_2", "context_embedder_2.linear_3", "time_embed.timestep_embedder.linear_1",
               "time_embed.timestep_embedder.linear_2", "proj_out", "x_embedder.proj"]
        lin += ["image_embedder.norm_in", "image_embedder.norm_out", "context_embedder_2.norm"]   # LayerNorm: weight + bias
        keys = ["cond_type_embed.weight"]
        for i in range(c.num_refiner_layers):
            q = f"context_embedder.token_refiner.refiner_blocks.{i}"
            lin += [q + ".norm1", q + ".norm2", q + ".ff.net.0.proj", q + ".ff.net.2", q + ".norm_out.linear"]
            lin += [f"{q}.attn.{n}" for n in ("to_q", "to_k", "to_v", "to_out.0")]
        lin += [n for n, _ in self._modulation_layout()]
        for i in range(c.num_layers):
            p = f"transformer_blocks.{i}"
            lin += [f"{p}.attn.{n}" for n in ("to_q", "to_k", "to_v", "to_out.0", "add_q_proj", "add_k_proj", "add_v_proj",
                                              "to_add_out")]
            lin += [f"{p}.{f}.net.0.proj" for f in ("ff", "ff_context")] + [f"{p}.{f}.net.2" for f in ("ff", "ff_context")]
            keys += [f"{p}.attn.{n}.weight" for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k")]
        for m in lin:
            keys += [m + ".weight", m + ".bias"]
        return keys

    def invalidate_caches(self) -> None:
        """Weights changed (load, LoRA merge): drop what was computed from them."""
        self._cond_plans.clear()

    def _finish_weights(self, w: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        self.invalidate_caches()
        c = self.config
        # Conv3d(kernel = stride = patch) as a GEMM over [tokens, C*pt*p*p]; K padded to a multiple of 8 for TMA (65 -> 72)
        pw = w["x_embedder.proj.weight"].reshape(c.inner_dim, -1)
        k = pw.shape[1]
        self._k_pad = (-k) % 8
        w["x_embedder.proj.weight"] = F.pad(pw, (0, self._k_pad)).contiguous()
        # output rows padded to a multiple of 8 (16-byte rows of the GEMM output); sliced off after the call
        self._n_out = w["proj_out.weight"].shape[0]
        n_pad = (-self._n_out) % 8
        if n_pad:
            w["proj_out.weight"] = F.pad(w["proj_out.weight"], (0, 0, 0, n_pad)).contiguous()
            w["proj_out.bias"] = F.pad(w["proj_out.bias"], (0, n_pad)).contiguous()
        for i in range(c.num_refiner_layers):
            a = f"context_embedder.token_refiner.refiner_blocks.{i}.attn"
            fuse_linears(w, a + ".to_qkv", [a + ".to_q", a + ".to_k", a + ".to_v"])
        for i in range(c.num_layers):
            a = f"transformer_blocks.{i}.attn"
            fuse_linears(w, a + ".to_qkv", [a + ".to_q", a + ".to_k", a + ".to_v"])
            fuse_linears(w, a + ".add_qkv", [a + ".add_q_proj", a + ".add_k_proj", a + ".add_v_proj"])
        r0 = 0
        lay = self._modulation_layout()
        for name, rows in lay:
            self._mod_rows[name] = (r0, rows)
            r0 += rows
        fuse_linears(w, "modulation", [n for n, _ in lay])
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
        """Synthetic weights of the architecture's shapes generated ON the device (bench; no checkpoints offline)."""
        dev = torch.device(device)
        self.device = dev
        g = torch.Generator(device=dev).manual_seed(seed)
        c, bf = self.config, torch.bfloat16
        d = c.inner_dim
        w: Dict[str, torch.Tensor] = {}

        def rnd(*shape, scale=std, base=0.0):
            return (torch.randn(*shape, generator=g, device=dev, dtype=torch.float32) * scale + base).to(bf)

        dims = self._linear_dims()
        for name, (out_f, in_f) in dims.items():
            w[name + ".weight"], w[name + ".bias"] = rnd(out_f, in_f), rnd(out_f)
        for name, n in self._norm_dims().items():
            w[name + ".weight"], w[name + ".bias"] = rnd(n, base=1.0), rnd(n)
        w["cond_type_embed.weight"] = rnd(3, d, scale=0.5)
        w["x_embedder.proj.weight"] = rnd(d, c.in_channels, c.patch_size_t, c.patch_size, c.patch_size)
        w["x_embedder.proj.bias"] = rnd(d)
        for i in range(c.num_layers):
            for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
                w[f"transformer_blocks.{i}.attn.{n}.weight"] = rnd(128, base=1.0)
        assert set(w) == set(self.state_dict_keys()), sorted(set(w) ^ set(self.state_dict_keys()))[:5]
        self.w = self._finish_weights(w)
        return self

    def _linear_dims(self) -> Dict[str, Tuple[int, int]]:
        c, d = self.config, self.config.inner_dim
        ffn = int(d * c.mlp_ratio)
        dims = {"image_embedder.linear_1": (c.image_embed_dim, c.image_embed_dim), "image_embedder.linear_2": (d, c.image_embed_dim),
                "context_embedder.proj_in": (d, c.text_embed_dim),
                "context_embedder.time_text_embed.timestep_embedder.linear_1": (d, 256),
                "context_embedder.time_text_embed.timestep_embedder.linear_2": (d, d),
                "context_embedder.time_text_embed.text_embedder.linear_1": (d, c.text_embed_dim),
                "context_embedder.time_text_embed.text_embedder.linear_2": (d, d),
                "context_embedder_2.linear_1": (2048, c.text_embed_2_dim), "context_embedder_2.linear_2": (2048, 2048),
                "context_embedder_2.linear_3": (d, 2048), "time_embed.timestep_embedder.linear_1": (d, 256),
                "time_embed.timestep_embedder.linear_2": (d, d),
                "proj_out": (c.patch_size_t * c.patch_size * c.patch_size * c.out_channels, d)}
        for i in range(c.num_refiner_layers):
            q = f"context_embedder.token_refiner.refiner_blocks.{i}"
            dims.update({q + ".ff.net.0.proj": (4 * d, d), q + ".ff.net.2": (d, 4 * d), q + ".norm_out.linear": (2 * d, d)})
            dims.update({f"{q}.attn.{n}": (d, d) for n in ("to_q", "to_k", "to_v", "to_out.0")})
        for name, rows in self._modulation_layout():
            dims[name] = (rows, d)
        for i in range(c.num_layers):
            p = f"transformer_blocks.{i}"
            dims.update({f"{p}.attn.{n}": (d, d) for n in ("to_q", "to_k", "to_v", "to_out.0", "add_q_proj", "add_k_proj",
                                                           "add_v_proj", "to_add_out")})
            for f in ("ff", "ff_context"):
                dims[f"{p}.{f}.net.0.proj"], dims[f"{p}.{f}.net.2"] = (ffn, d), (d, ffn)
        return dims

    def _norm_dims(self) -> Dict[str, int]:
        c, d = self.config, self.config.inner_dim
        n = {"image_embedder.norm_in": c.image_embed_dim, "image_embedder.norm_out": d, "context_embedder_2.norm": c.text_embed_2_dim}
        for i in range(c.num_refiner_layers):
            q = f"context_embedder.token_refiner.refiner_blocks.{i}"
            n[q + ".norm1"], n[q + ".norm2"] = d, d
        return n

    def lora_target(self, module: str):
        d = self.config.inner_dim
        head, _, leaf = module.rpartition(".")
        fused = None
        if head.endswith(".attn") and leaf in ("to_q", "to_k", "to_v"):
            fused = (head + ".to_qkv", ("to_q", "to_k", "to_v").index(leaf) * d, d)
        elif head.endswith(".attn") and leaf in ("add_q_proj", "add_k_proj", "add_v_proj"):
            fused = (head + ".add_qkv", ("add_q_proj", "add_k_proj", "add_v_proj").index(leaf) * d, d)
        elif module in self._mod_rows:
            fused = ("modulation",) + self._mod_rows[module]
        if fused is not None and fused[0] + ".weight" in self.w:
            return fused[0] + ".weight", fused[1], fused[2], fused[0] + ".bias"
        if module + ".weight" in self.w and self.w[module + ".weight"].dim() == 2 and module != "x_embedder.proj":
            return module + ".weight", 0, self.w[module + ".weight"].shape[0], module + ".bias"
        raise ValueError(f"Target module {module} not found in the model (or not a linear layer the b200 path adapts)")

    def parameter_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.w.values())

    # ------------------------------------------------------------------------------------ embedders
    def _lin(self, x: torch.Tensor, name: str, epilogue: int = ops.EPI_BIAS, **kw) -> torch.Tensor:
        return ops.linear(x, self.w[name + ".weight"], self.w.get(name + ".bias"), epilogue=epilogue, **kw)

    def _ln(self, x: torch.Tensor, name: str, eps: float) -> torch.Tensor:
        return ops.layernorm_modulate(x, ln_weight=self.w[name + ".weight"], ln_bias=self.w[name + ".bias"], eps=eps)

    def _timestep_mlp(self, t_proj: torch.Tensor, name: str) -> torch.Tensor:
        return self._lin(self._lin(t_proj, name + ".linear_1", ops.EPI_SILU), name + ".linear_2")

    def time_embed(self, timestep: torch.Tensor) -> torch.Tensor:
        """HunyuanVideo15TimeEmbedding (model.py:224-268): the sinusoid is cast to the timestep's (bf16) dtype."""
        return self._timestep_mlp(sinusoid_256(timestep, self.device).to(torch.bfloat16), "time_embed.timestep_embedder")

    def token_refiner(self, text_valid: torch.Tensor, timestep: torch.Tensor) -> torch.Tensor:
        """HunyuanVideo15TokenRefiner (model.py:450-496, blocks :297-412) on the VALID tokens [n, text_dim] of one sample."""
        c, w, p = self.config, self.w, "context_embedder."
        d, H = c.inner_dim, c.num_attention_heads
        n = text_valid.shape[0]
        pooled = text_valid.float().mean(dim=0, keepdim=True).to(torch.bfloat16)          # masked mean over valid tokens
        temb = (self._timestep_mlp(sinusoid_256(timestep, self.device).to(torch.bfloat16), p + "time_text_embed.timestep_embedder")
                + self._lin(self._lin(pooled, p + "time_text_embed.text_embedder.linear_1", ops.EPI_SILU),
                            p + "time_text_embed.text_embedder.linear_2"))
        act = F.silu(temb)
        h = self._lin(text_valid, p + "proj_in")
        qkv = torch.empty(n, 3 * d, dtype=torch.bfloat16, device=self.device)
        attn = torch.empty(n, d, dtype=torch.bfloat16, device=self.device)
        as4 = lambda t: t.view(1, n, H, 128).transpose(1, 2)
        for i in range(c.num_refiner_layers):
            q = p + f"token_refiner.refiner_blocks.{i}"
            gate_msa, gate_mlp = self._lin(act, q + ".norm_out.linear")[0].chunk(2)
            self._lin(self._ln(h, q + ".norm1", 1e-6), q + ".attn.to_qkv", out=qkv)
            ops.attention(as4(qkv[:, :d]), as4(qkv[:, d:2 * d]), as4(qkv[:, 2 * d:]), out=as4(attn))
            self._lin(attn, q + ".attn.to_out.0", ops.EPI_GATE_RES, out=h, gate=gate_msa)
            f1 = self._lin(self._ln(h, q + ".norm2", 1e-6), q + ".ff.net.0.proj", ops.EPI_SILU)
            self._lin(f1, q + ".ff.net.2", ops.EPI_GATE_RES, out=h, gate=gate_mlp)
        return h

    def _condition_plan(self, text, mask, text2, mask2, image_embeds) -> dict:
        """Everything of condition_tokens that does not depend on the timestep, computed ONCE per prompt: the valid counts and
        the t2v flag (host syncs), the compacted text tokens, and the image / byT5 branches.  Keyed on the identity and version
        of the five input tensors (the denoise loop passes the same device tensors every step); the entry keeps references to
        them, so a key cannot be recycled by another tensor while it is cached.  Without it every forward stalled the launch
        queue on four device -> host reads per sample and could not be captured in a CUDA graph."""
        key = tuple((t.data_ptr(), t._version, tuple(t.shape), t.dtype, str(t.device)) for t in (text, mask, text2, mask2, image_embeds))
        plan = self._cond_plans.get(key)
        if plan is not None:
            return plan
        w, d = self.w, self.config.inner_dim
        emb = w["cond_type_embed.weight"]
        v1, v2 = mask.bool().to(self.device), mask2.bool().to(self.device)
        n1, n2 = int(v1.sum()), int(v2.sum())
        plan = {"keep": (text, mask, text2, mask2, image_embeds), "n_pad": (v1.numel() - n1) + (v2.numel() - n2),
                "valid": [], "invalid": [], "text_valid": text[v1].contiguous() if n1 > 0 else None}
        if bool(torch.all(image_embeds == 0)):   # t2v: projection * 0 + type embedding, every token "invalid" (kept, not zeroed, :1036-1043)
            plan["invalid"].append(emb[2][None, :].expand(image_embeds.shape[0], d))
        else:
            h3 = self._ln(image_embeds, "image_embedder.norm_in", 1e-5)
            h3 = self._lin(self._lin(h3, "image_embedder.linear_1", ops.EPI_GELU_ERF), "image_embedder.linear_2")
            plan["valid"].append(self._ln(h3, "image_embedder.norm_out", 1e-5) + emb[2])
        if n2 > 0:
            h2 = self._ln(text2[v2].contiguous(), "context_embedder_2.norm", 1e-5)
            h2 = self._lin(self._lin(h2, "context_embedder_2.linear_1", ops.EPI_GELU_ERF), "context_embedder_2.linear_2",
                           ops.EPI_GELU_ERF)
            plan["valid"].append(self._lin(h2, "context_embedder_2.linear_3") + emb[1])
        plan["zeros"] = torch.zeros(plan["n_pad"], d, dtype=torch.bfloat16, device=self.device)
        if len(self._cond_plans) >= 8:           # cond / uncond of a few prompts; oldest first out
            self._cond_plans.pop(next(iter(self._cond_plans)))
        self._cond_plans[key] = plan
        return plan

    def condition_tokens(self, text, mask, text2, mask2, image_embeds, timestep) -> torch.Tensor:
        """model.py:1011-1101 for ONE sample: text [L1, text_dim], mask [L1], text2 [L2, text2_dim], mask2 [L2],
        image_embeds [L3, image_dim] -> [L3 + L2 + L1, dim] in the reference's valid-first order.  Only the token refiner
        (it is modulated by the timestep) runs per step; the rest comes from the per-prompt plan."""
        plan = self._condition_plan(text, mask, text2, mask2, image_embeds)
        parts = list(plan["valid"])
        if plan["text_valid"] is not None:
            parts.append(self.token_refiner(plan["text_valid"], timestep) + self.w["cond_type_embed.weight"][0])
        return torch.cat(parts + plan["invalid"] + [plan["zeros"]], dim=0)

    # ------------------------------------------------------------------------------------ forward
    def _rope(self, grid: Tuple[int, int, int]) -> torch.Tensor:
        key = (grid, str(self.device))
        if key not in self._rope_cache:
            self._rope_cache[key] = hy15_rope_table(grid, self.config.rope_axes_dim, self.config.rope_theta, self.device)
        return self._rope_cache[key]

    def patchify(self, x: torch.Tensor) -> torch.Tensor:
        """[C,F,H,W] -> [tokens, C*pt*p*p (+pad)] in the Conv3d's (C, pt, ph, pw) weight order."""
        c = self.config
        ch, f, h, w_ = x.shape
        pt, p = c.patch_size_t, c.patch_size
        t = x.reshape(ch, f // pt, pt, h // p, p, w_ // p, p).permute(1, 3, 5, 0, 2, 4, 6).reshape(-1, ch * pt * p * p)
        return F.pad(t, (0, self._k_pad)).contiguous() if self._k_pad else t.contiguous()

    @torch.inference_mode()
    def forward(self, hidden_states: torch.Tensor, timestep: torch.Tensor, encoder_hidden_states: torch.Tensor,
                encoder_attention_mask: torch.Tensor, timestep_r=None, encoder_hidden_states_2: Optional[torch.Tensor] = None,
                encoder_attention_mask_2: Optional[torch.Tensor] = None, image_embeds: Optional[torch.Tensor] = None,
                attention_kwargs=None, rope_on_cpu=None, return_dict: bool = False, **unused):
        if not self.w:
            raise RuntimeError("weights not loaded: call load_state_dict() or init_random_weights()")
        if timestep_r is not None:
            raise ValueError("timestep_r (meanflow) is not implemented on the b200 path")
        c, w, bf, dev = self.config, self.w, torch.bfloat16, self.device
        x_in = hidden_states.to(device=dev, dtype=bf)
        text, text2 = encoder_hidden_states.to(device=dev, dtype=bf), encoder_hidden_states_2.to(device=dev, dtype=bf)
        img = image_embeds.to(device=dev, dtype=bf)
        b, ch, f, hh, ww = x_in.shape
        pt, p = c.patch_size_t, c.patch_size
        if f % pt or hh % p or ww % p:
            raise ValueError(f"Input dims must be divisible by the patch size. Got (T,H,W)=({f},{hh},{ww})")
        grid = (f // pt, hh // p, ww // p)
        n_lat = grid[0] * grid[1] * grid[2]
        d, H = c.inner_dim, c.num_attention_heads
        from ..parallel import ParallelContext
        par: ParallelContext = unused.pop("parallel", None) or ParallelContext.single()
        lo, hi = par.shard_bounds(n_lat)              # this rank's latent-token shard (the whole sequence when sp_size == 1)
        n_loc = hi - lo
        rope = self._rope(grid)[lo:hi]
        t = timestep.to(device=dev, dtype=bf)                     # the engine passes it in the latent dtype (t2v.py:243)
        temb = self.time_embed(t)                                 # [B, dim]
        mod_all = ops.linear(F.silu(temb), w["modulation.weight"], w["modulation.bias"])
        outs = []
        for bi in range(b):
            ctx = self.condition_tokens(text[bi], encoder_attention_mask[bi], text2[bi], encoder_attention_mask_2[bi], img[bi],
                                        t[bi:bi + 1])
            n_ctx = ctx.shape[0]
            ws = self._ws
            if ws is None or ws.tokens != n_loc + n_ctx:
                self._ws = ws = JointWorkspace(n_loc + n_ctx, d, int(d * c.mlp_ratio), dev)
            ops.linear(self.patchify(x_in[bi])[lo:hi], w["x_embedder.proj.weight"], w["x_embedder.proj.bias"], out=ws.h[:n_loc])
            ws.h[n_loc:].copy_(ctx)
            m = mod_all[bi]
            for i in range(c.num_layers):
                pfx = f"transformer_blocks.{i}"

                def mods(name):
                    r0, rows = self._mod_rows[name]
                    return m[r0:r0 + rows].chunk(6)

                streams = (
                    StreamParams(slice(0, n_loc), mods(pfx + ".norm1.linear"), pfx + ".attn.to_qkv", pfx + ".attn.norm_q.weight",
                                 pfx + ".attn.norm_k.weight", pfx + ".attn.to_out.0", pfx + ".ff", rope),
                    StreamParams(slice(n_loc, None), mods(pfx + ".norm1_context.linear"), pfx + ".attn.add_qkv",
                                 pfx + ".attn.norm_added_q.weight", pfx + ".attn.norm_added_k.weight", pfx + ".attn.to_add_out",
                                 pfx + ".ff_context", None),
                )
                dual_stream_block(w, ws, streams, H, ops.NORM_INPLACE_RMS, par=par, n_img_total=n_lat)
            r0, rows = self._mod_rows["norm_out.linear"]
            scale, shift = m[r0:r0 + rows].chunk(2)                # AdaLayerNormContinuous: scale first
            ops.adaln_zero_modulate(ws.h[:n_loc], scale, shift, out=ws.norm[:n_loc])
            y_loc = ops.linear(ws.norm[:n_loc], w["proj_out.weight"], w["proj_out.bias"])[:, :self._n_out]
            outs.append(par.gather_tokens(y_loc.contiguous()))
        y = torch.stack(outs, dim=0).reshape(b, grid[0], grid[1], grid[2], -1, pt, p, p).permute(0, 4, 1, 5, 2, 6, 3, 7)
        out = y.flatten(6, 7).flatten(4, 5).flatten(2, 3)
        if return_dict:
            return {"sample": out}
        return (out,)

    __call__ = forward

    def set_chunking_profile(self, profile_name: str) -> None:
        """Chunking profiles exist to fit small GPUs (model.py:758-781, :905-929); on 180 GB they are pure overhead."""
        if profile_name not in ("none", "light", "balanced", "aggressive"):
            raise ValueError(f"Unknown chunking profile '{profile_name}'. Available: ['aggressive', 'balanced', 'light', 'none']")

    def eval(self):
        return self

    def to(self, *args, **kwargs):
        return self
