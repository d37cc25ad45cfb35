"""LoRA hooks of the fused path: adapters are MERGED into the resident bf16 weights by the tcgen05 GEMM.

Reference behaviour (apps/api/src/lora/manager.py:454-606, applied from engine/base_engine.py:1303-1317): the manager
calls ``model.load_lora_adapter(state_dict, adapter_name=..., prefix=..., metadata=...)`` once per file and then
``model.set_adapters(names, weights=scales)``; PEFT injects ``lora_A`` / ``lora_B`` modules, so every adapted
``nn.Linear`` costs two extra skinny GEMMs, an elementwise scale and an add per call, each rounding to bf16:

    y = W x + b + scaling * B (A x)            scaling = weight * alpha_m / r_m

with ``alpha_m = alpha_pattern.get(m, lora_alpha)`` and ``r_m = rank_pattern.get(m, r)`` per target module m (PEFT
``LoraLayer.update_layer``).  The manager's metadata (manager.py:433-447) sets ``r = lora_alpha =`` the MOST COMMON rank of the
file and lists the other ranks in ``rank_pattern``, so a module whose rank differs from the modal rank is scaled by
``r_modal / r_m`` -- reproduced here (``peft_scaling``), from the metadata when it is passed, else recomputed the way the
manager computes it.

Here the adapters never touch the per-step path.  ``set_adapters`` rebuilds the effective weight of every adapted
projection once, on the device, with ONE launch of ``b200_linear`` per projection:

    W_eff[N, K]  =  W_base[N, K]  +  [s1*B1 | s2*B2 | ...][N, R]  @  [A1; A2; ...][R, K]       (R = sum of ranks)

computed as ``linear(x = B_cat, weight = A_cat^T, epilogue = GATE_RES, out = W_eff)``: fp32 accumulation in TMEM, added
to the fp32 value of the base weight in the epilogue, ONE rounding to bf16 (PEFT's own ``merge`` rounds the delta and
the sum separately).  The fused projections of this package (``attn1.to_qkv``, ``attn2.to_kv``) are addressed by row
slice, so ``to_q`` / ``to_k`` / ``to_v`` adapters land in the right rows.  The pristine rows of every adapted weight
are kept (device by default -- 180 GB of HBM -- or pinned host memory), so changing scales or removing adapters is
exact: ``set_adapters([], [])`` restores the base weights bit for bit.

The mixin mirrors the part of diffusers' ``PeftAdapterMixin`` the manager relies on: ``peft_config`` (dict keyed by
adapter name), ``load_lora_adapter``, ``set_adapters``, ``delete_adapters``, ``disable_lora`` / ``enable_lora``.
State-dict normalisation follows manager.py (``_strip_adapter_name_from_keys`` :812-838, ``_get_prefix_key``
:383-396) and lora_converter.py (``scale_alpha`` :139-163, ``lora_down/lora_up`` renaming); the format zoo of the
converter (Kohya, old diffusers) is loader territory and out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import torch

from . import ops

KNOWN_PREFIXES = ("transformer", "diffusion_model", "model", "unet")


@dataclass
class LoraAdapter:
    """One loaded adapter: per target module the bf16 factors A [r, K] and B [N, r] (+ optional lora_B bias)."""
    name: str
    A: Dict[str, torch.Tensor] = field(default_factory=dict)
    B: Dict[str, torch.Tensor] = field(default_factory=dict)
    bias: Dict[str, torch.Tensor] = field(default_factory=dict)
    rank: Dict[str, int] = field(default_factory=dict)
    scaling: Dict[str, float] = field(default_factory=dict)   # PEFT's alpha_m / r_m per module (1.0 at the modal rank)

    @property
    def modules(self) -> List[str]:
        return sorted(self.A)


# ---------------------------------------------------------------------------------------------------------
# state-dict normalisation (host logic, no GPU)
# ---------------------------------------------------------------------------------------------------------
def strip_adapter_name_from_keys(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """``x.lora_B.default.weight`` -> ``x.lora_B.weight`` (manager.py:812-838)."""
    out = {}
    for key, value in state.items():
        parts = key.split(".")
        if (len(parts) >= 3 and parts[-3] in ("lora_A", "lora_B") and parts[-1] in ("weight", "bias", "alpha")
                and parts[-2] not in ("lora_A", "lora_B")):
            parts.pop(-2)
            key = ".".join(parts)
        out[key] = value
    return out


def get_prefix_key(keys: Sequence[str]) -> Optional[str]:
    """manager.py:383-396: a prefix counts only when both the first and the last key carry it."""
    for p in KNOWN_PREFIXES:
        if keys and keys[0].startswith(p + ".") and keys[-1].startswith(p + "."):
            return p
    return None


def scale_alpha(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Fold ``<module>.alpha`` into the factors like lora_converter.py:139-163 (alpha / rank, split between down and
    up by powers of two) and drop the alpha entries (PEFT weights carry none; the manager sets lora_alpha = r)."""
    out = dict(state)
    for key in list(state):
        if not key.endswith(".alpha"):
            continue
        down_key, up_key = key[:-len(".alpha")] + ".lora_A.weight", key[:-len(".alpha")] + ".lora_B.weight"
        if down_key in state and up_key in state:
            rank = state[down_key].shape[0]
            scale_down, scale_up = float(state[key].item()) / rank, 1.0
            while scale_down * 2 < scale_up:
                scale_down *= 2
                scale_up /= 2
            out[down_key] = state[down_key] * scale_down
            out[up_key] = state[up_key] * scale_up
        out.pop(key)
    return out


def normalize_lora_state_dict(state: Dict[str, torch.Tensor], prefix: Optional[str] = "auto") -> Dict[str, torch.Tensor]:
    """-> ``{<module>.lora_A.weight | <module>.lora_B.weight | <module>.lora_B.bias: tensor}`` with the model prefix,
    the adapter-name segment and alpha entries removed and ``lora_down/lora_up`` renamed."""
    st = {k.replace(".lora_down.", ".lora_A.").replace(".lora_up.", ".lora_B."): v for k, v in state.items()}
    st = strip_adapter_name_from_keys(st)
    st = scale_alpha(st)
    keys = list(st)
    if prefix == "auto":
        prefix = get_prefix_key(keys)
    if prefix:
        # diffusers keeps only the keys under the prefix and strips it
        st = {k[len(prefix) + 1:]: v for k, v in st.items() if k.startswith(prefix + ".")}
    bad = [k for k in st if "lora_magnitude_vector" in k or "lora_embedding" in k]
    if bad:
        raise ValueError(f"DoRA / embedding LoRA keys are not supported on the b200 path: {bad[:3]}")
    return st


def split_modules(state: Dict[str, torch.Tensor]) -> Dict[str, Dict[str, torch.Tensor]]:
    """group a normalised state dict by target module -> {"A": [r,K], "B": [N,r], "bias": [N]?}"""
    mods: Dict[str, Dict[str, torch.Tensor]] = {}
    for k, v in state.items():
        for suffix, slot in ((".lora_A.weight", "A"), (".lora_B.weight", "B"), (".lora_B.bias", "bias")):
            if k.endswith(suffix):
                mods.setdefault(k[:-len(suffix)], {})[slot] = v
                break
        else:
            raise ValueError(f"unrecognised LoRA key: {k}")
    for m, d in mods.items():
        if "A" not in d or "B" not in d:
            raise ValueError(f"LoRA module {m} needs both lora_A.weight and lora_B.weight")
        if d["A"].dim() != 2 or d["B"].dim() != 2 or d["A"].shape[0] != d["B"].shape[1]:
            raise ValueError(f"LoRA module {m}: A {tuple(d['A'].shape)} / B {tuple(d['B'].shape)} are not [r,K] / [N,r]")
    return mods


def _pattern_get(pattern: Optional[dict], module: str, default):
    """PEFT looks a module up in ``rank_pattern`` / ``alpha_pattern`` by exact name or by name suffix."""
    if not pattern:
        return default
    if module in pattern:
        return pattern[module]
    for key, val in pattern.items():
        if module.endswith("." + key):
            return val
    return default


def peft_scaling(ranks: Dict[str, int], metadata: Optional[dict] = None) -> Tuple[Dict[str, float], dict]:
    """Per-module ``alpha_m / r_m`` as PEFT computes it from the LoraConfig the manager builds (manager.py:433-447):
    without explicit metadata ``r = lora_alpha =`` the most common rank (first seen wins a tie, as
    ``collections.Counter.most_common`` does) and ``rank_pattern`` holds the other ranks.  Returns (scaling, config)."""
    import collections

    md = dict(metadata or {})
    if "r" not in md or "lora_alpha" not in md:
        r_modal = collections.Counter(ranks.values()).most_common(1)[0][0]
        md.setdefault("r", int(r_modal))
        md.setdefault("lora_alpha", int(md["r"]))
        md.setdefault("rank_pattern", {m: rr for m, rr in ranks.items() if rr != md["r"]})
        md.setdefault("alpha_pattern", {})
    scaling = {}
    for m, r_actual in ranks.items():
        r_m = int(_pattern_get(md.get("rank_pattern"), m, md["r"]))
        if r_m != r_actual:
            raise ValueError(f"LoRA module {m}: metadata says rank {r_m}, the factors have rank {r_actual}")
        alpha_m = float(_pattern_get(md.get("alpha_pattern"), m, md["lora_alpha"]))
        scaling[m] = alpha_m / r_m
    return scaling, md


# ---------------------------------------------------------------------------------------------------------
# the mixin
# ---------------------------------------------------------------------------------------------------------
class LoraHostMixin:
    """Needs ``self.w`` (flat dict of bf16 CUDA tensors) and ``self.lora_target(module) -> (wkey, row0, rows, bkey)``."""

    #: where the pristine rows of adapted weights are kept: "cuda" (default) or "cpu" (pinned)
    lora_base_device = "cuda"

    def _lora_state(self):
        if not hasattr(self, "_lora_adapters"):
            self._lora_adapters: Dict[str, LoraAdapter] = {}
            self._lora_scales: Dict[str, float] = {}
            self._lora_base: Dict[str, Tuple[torch.Tensor, Optional[torch.Tensor]]] = {}
            self._lora_disabled = False
            self.peft_config: Dict[str, dict] = {}
        return self._lora_adapters

    # -- PeftAdapterMixin surface ---------------------------------------------------------------------------
    def load_lora_adapter(self, state_dict: Dict[str, torch.Tensor], adapter_name: str = "default",
                          prefix: Optional[str] = "transformer", metadata: Optional[dict] = None, **unused) -> None:
        """Register an adapter (no weights change until ``set_adapters``).  ``prefix`` as in diffusers: keys under
        ``<prefix>.`` are kept and stripped; ``None`` = keys are already relative to the model."""
        adapters = self._lora_state()
        if adapter_name in adapters:
            raise ValueError(f"Adapter name {adapter_name} already in use in the model - please select a new adapter name.")
        if not self.w:
            raise RuntimeError("weights not loaded: call load_state_dict() before loading a LoRA")
        keys = list(state_dict)
        if prefix is not None and not any(k.startswith(prefix + ".") for k in keys):
            prefix = None   # diffusers warns and loads nothing; the manager only passes a prefix it has just detected
        st = normalize_lora_state_dict(state_dict, prefix=prefix)
        ad = LoraAdapter(adapter_name)
        for module, d in split_modules(st).items():
            wkey, row0, rows, _ = self.lora_target(module)
            K = self.w[wkey].shape[1]
            if d["B"].shape[0] != rows or d["A"].shape[1] != K:
                raise ValueError(f"LoRA module {module}: factors {tuple(d['B'].shape)} x {tuple(d['A'].shape)} do not match "
                                 f"the weight [{rows}, {K}]")
            dev = self.w[wkey].device
            ad.A[module] = d["A"].detach().to(device=dev, dtype=torch.float32)
            ad.B[module] = d["B"].detach().to(device=dev, dtype=torch.float32)
            if "bias" in d:
                ad.bias[module] = d["bias"].detach().to(device=dev, dtype=torch.float32)
            ad.rank[module] = int(d["A"].shape[0])
        if not ad.A:
            raise ValueError("no LoRA weights found for this model in the state dict")
        ad.scaling, cfg = peft_scaling(ad.rank, metadata)
        adapters[adapter_name] = ad
        self.peft_config[adapter_name] = dict(cfg, target_modules=ad.modules)
        self._lora_scales.setdefault(adapter_name, 1.0)

    def set_adapters(self, adapter_names: Union[str, Sequence[str]],
                     weights: Optional[Union[float, Sequence[float]]] = None) -> None:
        """Activate ``adapter_names`` with ``weights`` (default 1.0) and deactivate every other adapter, then rebuild
        the effective weights of every projection any adapter touches."""
        adapters = self._lora_state()
        names = [adapter_names] if isinstance(adapter_names, str) else list(adapter_names)
        if weights is None:
            weights = [1.0] * len(names)
        elif isinstance(weights, (int, float)):
            weights = [float(weights)] * len(names)
        if len(weights) != len(names):
            raise ValueError(f"Length of adapter names {len(names)} is not equal to the length of their weights {len(weights)}.")
        for n in names:
            if n not in adapters:
                raise ValueError(f"Adapter {n} is not loaded (loaded: {sorted(adapters)})")
        self._lora_scales = {n: 0.0 for n in adapters}
        for n, s in zip(names, weights):
            self._lora_scales[n] = 1.0 if s is None else float(s)
        self._remerge()

    def delete_adapters(self, adapter_names: Union[str, Sequence[str]]) -> None:
        adapters = self._lora_state()
        for n in ([adapter_names] if isinstance(adapter_names, str) else list(adapter_names)):
            if n not in adapters:
                raise ValueError(f"Adapter name {n} not found in the model")
            touched = adapters[n].modules
            del adapters[n]
            self._lora_scales.pop(n, None)
            self.peft_config.pop(n, None)
            self._remerge(touched)

    def disable_lora(self) -> None:
        self._lora_state()
        self._lora_disabled = True
        self._remerge()

    def enable_lora(self) -> None:
        self._lora_state()
        self._lora_disabled = False
        self._remerge()

    def active_adapters(self) -> List[str]:
        self._lora_state()
        return [] if self._lora_disabled else [n for n, s in self._lora_scales.items() if s != 0.0]

    # -- merge ------------------------------------------------------------------------------------------------
    def _keep_base(self, module: str) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        if module not in self._lora_base:
            wkey, row0, rows, bkey = self.lora_target(module)

            def keep(t: torch.Tensor) -> torch.Tensor:
                if self.lora_base_device == "cpu":
                    return t.detach().to("cpu").pin_memory()
                return t.detach().clone()

            b = keep(self.w[bkey][row0:row0 + rows]) if (bkey is not None and bkey in self.w) else None
            self._lora_base[module] = (keep(self.w[wkey][row0:row0 + rows]), b)
        return self._lora_base[module]

    def _remerge(self, modules: Optional[Iterable[str]] = None) -> None:
        getattr(self, "invalidate_caches", lambda: None)()   # models may cache results computed from the weights
        adapters = self._lora_state()
        if modules is None:
            modules = sorted(set(self._lora_base) | {m for ad in adapters.values() for m in ad.modules})
        for module in modules:
            wkey, row0, rows, bkey = self.lora_target(module)
            w_rows = self.w[wkey][row0:row0 + rows]
            base_w, base_b = self._keep_base(module)
            w_rows.copy_(base_w, non_blocking=True)
            if base_b is not None:
                self.w[bkey][row0:row0 + rows].copy_(base_b, non_blocking=True)
            parts = [] if self._lora_disabled else [(ad, self._lora_scales.get(n, 0.0) * ad.scaling.get(module, 1.0))
                                                    for n, ad in adapters.items()
                                                    if module in ad.A and self._lora_scales.get(n, 0.0) != 0.0]
            if not parts:
                continue
            # [s1*B1 | s2*B2 | ...] in fp32, ONE rounding to bf16; rank padded to the 16-byte row pitch TMA needs
            b_cat = torch.cat([ad.B[module] * s for ad, s in parts], dim=1)
            a_cat = torch.cat([ad.A[module] for ad, _ in parts], dim=0)
            pad = (-b_cat.shape[1]) % 8
            if pad:
                b_cat = torch.nn.functional.pad(b_cat, (0, pad))
                a_cat = torch.nn.functional.pad(a_cat, (0, 0, 0, pad))
            ops.linear(b_cat.to(torch.bfloat16).contiguous(), a_cat.to(torch.bfloat16).t().contiguous(), None,
                       epilogue=ops.EPI_GATE_RES, out=w_rows, gate=None)
            bias_parts = [(ad.bias[module], s) for ad, s in parts if module in ad.bias]
            if bias_parts:
                if bkey is None or bkey not in self.w:
                    raise ValueError(f"LoRA module {module} carries a lora_B bias but the layer has no bias")
                acc = self.w[bkey][row0:row0 + rows].float()
                for bvec, s in bias_parts:
                    acc = acc + s * bvec
                self.w[bkey][row0:row0 + rows].copy_(acc.to(torch.bfloat16))
