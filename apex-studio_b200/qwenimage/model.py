"""QwenImage DiT forward on the B200 kernels -- BASELINE.json configs[2] (QwenImage-Edit-2509, 8-step Lightning).

Host-side mirror of the reference's ``QwenImageTransformer2DModel``
(apps/api/src/transformer/qwenimage/base/model.py:753, forward :853-993; block :679-750; attention processor :480-578;
RoPE :186-300): same constructor config names, same diffusers-format state-dict keys, same
``forward(hidden_states, encoder_hidden_states, encoder_hidden_states_mask, timestep, img_shapes, txt_seq_lens,
return_dict=False) -> (Tensor,)`` contract as ``engine/qwenimage/shared.py:394-404`` calls it (the edit engines append the
reference-image tokens to ``hidden_states`` and list their shapes in ``img_shapes``, edit_plus.py).

60 dual-stream blocks through ``mmdit.dual_stream_block``: one residual stream [S_txt + S_img, 3072] with the text rows
first (the reference concatenates [text, image], :552-556), per-head diffusers-RMSNorm + complex RoPE on BOTH streams
(``ops.NORM_DIFFUSERS_RMS``), the 60 x 2 modulation linears + the head's AdaLayerNormContinuous as ONE weight-streaming
GEMM per forward, the text-stream input RMSNorm as a row kernel.  The joint attention takes no mask (the processor passes
``attention_mask=None`` through, :558-565; ``encoder_hidden_states_mask`` is accepted and ignored exactly like the reference).
``zero_cond_t`` / ``use_additional_t_cond`` / ``use_layer3d_rope`` / ControlNet residuals are not implemented (raise).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from .. import ops
from ..lora import LoraHostMixin
from ..mmdit import JointWorkspace, StreamParams, dual_stream_block, fuse_linears, sinusoid_256


@dataclass
class QwenImageConfig:
    """Constructor arguments of the reference class (model.py:791-805); defaults = Qwen-Image / Qwen-Image-Edit-2509."""
    patch_size: int = 2
    in_channels: int = 64
    out_channels: Optional[int] = 16
    num_layers: int = 60
    attention_head_dim: int = 128
    num_attention_heads: int = 24
    joint_attention_dim: int = 3584
    guidance_embeds: bool = False
    axes_dims_rope: Tuple[int, int, int] = (16, 56, 56)
    zero_cond_t: bool = False
    use_additional_t_cond: bool = False
    use_layer3d_rope: bool = False

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim


def _rope_params(index: torch.Tensor, dim: int, theta: float) -> torch.Tensor:
    ang = torch.outer(index, 1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float32).div(dim)))
    return torch.view_as_real(torch.polar(torch.ones_like(ang), ang))     # (cos, sin) exactly as the reference builds them


def qwen_rope_tables(img_shapes: Sequence[Tuple[int, int, int]], txt_len: int, axes_dim, device, theta: float = 10000.0):
    """QwenEmbedRope.forward with scale_rope=True (model.py:229-300) for one sample -> fp32 (cos, sin) tables
    (img [S_img, 64, 2], txt [txt_len, 64, 2]): frame position = image index, height / width positions centred, text
    positions start at max(height/2, width/2) over all images."""
    pos_index = torch.arange(4096)
    neg_index = torch.arange(4096).flip(0) * -1 - 1
    pos = [_rope_params(pos_index, d, theta) for d in axes_dim]
    neg = [_rope_params(neg_index, d, theta) for d in axes_dim]
    vid, max_vid_index = [], 0
    for idx, (frame, height, width) in enumerate(img_shapes):
        ff = pos[0][idx:idx + frame].view(frame, 1, 1, -1, 2).expand(frame, height, width, -1, 2)
        fh = torch.cat([neg[1][-(height - height // 2):], pos[1][:height // 2]], dim=0).view(1, height, 1, -1, 2).expand(frame, height, width, -1, 2)
        fw = torch.cat([neg[2][-(width - width // 2):], pos[2][:width // 2]], dim=0).view(1, 1, width, -1, 2).expand(frame, height, width, -1, 2)
        vid.append(torch.cat([ff, fh, fw], dim=3).reshape(frame * height * width, -1, 2))
        max_vid_index = max(height // 2, width // 2, max_vid_index)
    txt = torch.cat(pos, dim=1)[max_vid_index:max_vid_index + txt_len]
    return torch.cat(vid, dim=0).contiguous().to(device), txt.contiguous().to(device)


class QwenImageTransformer2DModel(LoraHostMixin):
    """B200 implementation; see module docstring."""

    def __init__(self, config: Optional[QwenImageConfig] = None, **kwargs):
        self.config = config or QwenImageConfig(**kwargs)
        c = self.config
        if c.attention_head_dim != 128:
            raise ValueError("the b200 attention kernel supports attention_head_dim == 128 only")
        if sum(c.axes_dims_rope) != c.attention_head_dim:
            raise ValueError(f"axes_dims_rope {c.axes_dims_rope} must sum to attention_head_dim")
        for flag in ("zero_cond_t", "use_additional_t_cond", "use_layer3d_rope", "guidance_embeds"):
            if getattr(c, flag):
                raise ValueError(f"{flag}=True is not implemented on the b200 path")
        self.w: Dict[str, torch.Tensor] = {}
        self._mod_rows: Dict[str, Tuple[int, int]] = {}
        self._rope_cache: Dict[Tuple, Tuple[torch.Tensor, torch.Tensor]] = {}
        self._ws: Optional[JointWorkspace] = None
        self._n_out = 0
        self.dtype = torch.bfloat16
        self.device = None

    @classmethod
    def from_config(cls, config, **kwargs):
        if isinstance(config, QwenImageConfig):
            return cls(config)
        names = QwenImageConfig.__dataclass_fields__.keys()
        return cls(QwenImageConfig(**{k: (tuple(v) if k == "axes_dims_rope" else v) for k, v in dict(config).items()
                                      if k in names}), **kwargs)

    # ------------------------------------------------------------------------------------ weights
    def _modulation_layout(self) -> List[Tuple[str, int]]:
        c, d = self.config, self.config.inner_dim
        lay = []
        for i in range(c.num_layers):
            lay += [(f"transformer_blocks.{i}.img_mod.1", 6 * d), (f"transformer_blocks.{i}.txt_mod.1", 6 * d)]
        lay.append(("norm_out.linear", 2 * d))
        return lay

    def _linear_dims(self) -> Dict[str, Tuple[int, int]]:
        c, d = self.config, self.config.inner_dim
        dims = {"time_text_embed.timestep_embedder.linear_1": (d, 256), "time_text_embed.timestep_embedder.linear_2": (d, d),
                "img_in": (d, c.in_channels), "txt_in": (d, c.joint_attention_dim),
                "proj_out": (c.patch_size * c.patch_size * (c.out_channels or c.in_channels), d)}
        for name, rows in self._modulation_layout():
            dims[name] = (rows, d)
        for i in range(c.num_layers):
            p = f"transformer_blocks.{i}"
            dims.update({f"{p}.attn.{n}": (d, d) for n in ("to_q", "to_k", "to_v", "to_out.0", "add_q_proj", "add_k_proj",
                                                           "add_v_proj", "to_add_out")})
            for f in ("img_mlp", "txt_mlp"):
                dims[f"{p}.{f}.net.0.proj"], dims[f"{p}.{f}.net.2"] = (4 * d, d), (d, 4 * d)
        return dims

    def state_dict_keys(self) -> List[str]:
        keys = ["txt_norm.weight"]
        for m in self._linear_dims():
            keys += [m + ".weight", m + ".bias"]
        for i in range(self.config.num_layers):
            keys += [f"transformer_blocks.{i}.attn.{n}.weight" for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k")]
        return keys

    def _finish_weights(self, w: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        c = self.config
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
        self._n_out = w["proj_out.weight"].shape[0]
        n_pad = (-self._n_out) % 8     # GEMM output rows are 16-byte multiples; the pad columns are sliced off
        if n_pad:
            w["proj_out.weight"] = F.pad(w["proj_out.weight"], (0, 0, 0, n_pad)).contiguous()
            w["proj_out.bias"] = F.pad(w["proj_out.bias"], (0, n_pad)).contiguous()
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
        bf = torch.bfloat16

        def rnd(*shape, scale=std, base=0.0):
            return (torch.randn(*shape, generator=g, device=dev, dtype=torch.float32) * scale + base).to(bf)

        w: Dict[str, torch.Tensor] = {"txt_norm.weight": rnd(self.config.joint_attention_dim, base=1.0)}
        for name, (out_f, in_f) in self._linear_dims().items():
            w[name + ".weight"], w[name + ".bias"] = rnd(out_f, in_f), rnd(out_f)
        for i in range(self.config.num_layers):
            for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
                w[f"transformer_blocks.{i}.attn.{n}.weight"] = rnd(128, base=1.0)
        self.w = self._finish_weights(w)
        return self

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
        if module == "proj_out":
            return "proj_out.weight", 0, self._n_out, "proj_out.bias"
        if module + ".weight" in self.w and self.w[module + ".weight"].dim() == 2:
            return module + ".weight", 0, self.w[module + ".weight"].shape[0], module + ".bias"
        raise ValueError(f"Target module {module} not found in the model (or not a linear layer the b200 path adapts)")

    def parameter_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.w.values())

    # ------------------------------------------------------------------------------------ forward
    def _rope(self, img_shapes, txt_len: int):
        key = (tuple(tuple(int(v) for v in s) for s in img_shapes), int(txt_len), str(self.device))
        if key not in self._rope_cache:
            if len(self._rope_cache) > 8:
                self._rope_cache.clear()
            self._rope_cache[key] = qwen_rope_tables(key[0], txt_len, self.config.axes_dims_rope, self.device)
        return self._rope_cache[key]

    def time_embed(self, timestep: torch.Tensor) -> torch.Tensor:
        """QwenTimestepProjEmbeddings (model.py:154-183): Timesteps(scale=1000) sinusoid cast to bf16 -> MLP."""
        w, p = self.w, "time_text_embed.timestep_embedder"
        half = 128
        import math
        freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=self.device) / half)
        arg = 1000 * (timestep[:, None].float() * freqs[None, :])
        proj = torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1).to(torch.bfloat16)
        h1 = ops.linear(proj, w[p + ".linear_1.weight"], w[p + ".linear_1.bias"], epilogue=ops.EPI_SILU)
        return ops.linear(h1, w[p + ".linear_2.weight"], w[p + ".linear_2.bias"])

    @torch.inference_mode()
    def forward(self, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor = None,
                encoder_hidden_states_mask: torch.Tensor = None, timestep: torch.Tensor = None, img_shapes=None,
                txt_seq_lens: Optional[List[int]] = None, guidance=None, attention_kwargs=None, controlnet_block_samples=None,
                additional_t_cond=None, return_dict: bool = False, **unused):
        if not self.w:
            raise RuntimeError("weights not loaded: call load_state_dict() or init_random_weights()")
        if controlnet_block_samples is not None or additional_t_cond is not None:
            raise ValueError("ControlNet residuals / additional_t_cond are not implemented on the b200 path")
        c, w, bf, dev = self.config, self.w, torch.bfloat16, self.device
        x_in = hidden_states.to(device=dev, dtype=bf)
        enc = encoder_hidden_states.to(device=dev, dtype=bf)
        b, n_img, _ = x_in.shape
        n_txt = enc.shape[1]
        d, H = c.inner_dim, c.num_attention_heads
        if isinstance(img_shapes[0], (list, tuple)) and isinstance(img_shapes[0][0], (list, tuple)):
            shapes = img_shapes[0]          # the reference uses the first sample's shapes for the whole batch (:243-244)
        else:
            shapes = img_shapes if isinstance(img_shapes[0], (list, tuple)) else [img_shapes]
        if sum(int(f) * int(h) * int(w_) for f, h, w_ in shapes) != n_img:
            raise ValueError(f"img_shapes {shapes} do not add up to the {n_img} image tokens")
        txt_len = max(txt_seq_lens) if txt_seq_lens else n_txt
        if txt_len != n_txt:
            raise ValueError(f"max(txt_seq_lens) = {txt_len} must equal the text sequence length {n_txt}")
        from ..parallel import ParallelContext
        par: ParallelContext = unused.pop("parallel", None) or ParallelContext.single()
        lo, hi = par.shard_bounds(n_img)              # this rank's image-token shard (everything when sp_size == 1)
        n_loc = hi - lo
        img_rope, txt_rope = self._rope(shapes, txt_len)
        img_rope = img_rope[lo:hi]
        temb = self.time_embed(timestep.to(device=dev, dtype=bf))
        mod_all = ops.linear(F.silu(temb), w["modulation.weight"], w["modulation.bias"])
        ws = self._ws
        if ws is None or ws.tokens != n_txt + n_loc:
            self._ws = ws = JointWorkspace(n_txt + n_loc, d, 4 * d, dev)
        outs = []
        for bi in range(b):
            ops.linear(ops.rmsnorm_rows(enc[bi], w["txt_norm.weight"], 1e-6, ops.NORM_DIFFUSERS_RMS), w["txt_in.weight"],
                       w["txt_in.bias"], out=ws.h[:n_txt])
            ops.linear(x_in[bi, lo:hi], w["img_in.weight"], w["img_in.bias"], out=ws.h[n_txt:])
            m = mod_all[bi]

            def mods(name):
                r0, rows = self._mod_rows[name]
                return m[r0:r0 + rows].chunk(6)

            for i in range(c.num_layers):
                p = f"transformer_blocks.{i}"
                streams = (
                    StreamParams(slice(n_txt, None), mods(p + ".img_mod.1"), p + ".attn.to_qkv", p + ".attn.norm_q.weight",
                                 p + ".attn.norm_k.weight", p + ".attn.to_out.0", p + ".img_mlp", img_rope),
                    StreamParams(slice(0, n_txt), mods(p + ".txt_mod.1"), p + ".attn.add_qkv", p + ".attn.norm_added_q.weight",
                                 p + ".attn.norm_added_k.weight", p + ".attn.to_add_out", p + ".txt_mlp", txt_rope),
                )
                dual_stream_block(w, ws, streams, H, ops.NORM_DIFFUSERS_RMS, par=par, n_img_total=n_img)
            r0, rows = self._mod_rows["norm_out.linear"]
            scale, shift = m[r0:r0 + rows].chunk(2)
            ops.adaln_zero_modulate(ws.h[n_txt:], scale, shift, out=ws.norm[n_txt:])
            y_loc = ops.linear(ws.norm[n_txt:], w["proj_out.weight"], w["proj_out.bias"])[:, :self._n_out]
            outs.append(par.gather_tokens(y_loc.contiguous()))
        out = torch.stack(outs, dim=0)
        if return_dict:
            return {"sample": out}
        return (out,)

    __call__ = forward

    def eval(self):
        return self

    def to(self, *args, **kwargs):
        return self
