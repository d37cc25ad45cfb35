"""CPU oracle of the LoRA path -- TEST INFRASTRUCTURE (imported only by tests/).

The reference applies LoRAs through PEFT (apps/api/src/lora/manager.py:566-588: ``model.load_lora_adapter`` then
``model.set_adapters(names, weights=scales)``); ``peft`` and ``diffusers`` are un-vendored, unpinned dependencies
(api/requirements/requirements.txt) that are absent from /root/reference and from this image, so this file restates
their published arithmetic:

  peft.tuners.lora.layer.Linear.forward :  result = base_layer(x);  result = result + lora_B(lora_A(dropout(x))) * scaling
  diffusers set_adapters -> LoraLayer.set_scale :  scaling[adapter] = weight * alpha_m / r_m   (per module m:
      r_m = rank_pattern.get(m, r), alpha_m = alpha_pattern.get(m, lora_alpha) -- peft LoraLayer.update_layer)
  manager.py:433-447 :  r = lora_alpha = the MOST COMMON rank of the file, rank_pattern = the other ranks, so
      scaling_m = weight * r_modal / r_m   (= weight for single-rank files)        -> :func:`manager_scaling`

PARITY UNPINNED: the reference holds no test or golden vector for LoRA numerics (tests/ only check LoRA *resolution*).
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch
import torch.nn.functional as F

Weights = Dict[str, torch.Tensor]


def manager_scaling(ranks: Dict[str, int]) -> Dict[str, float]:
    """manager.py:433-447 + peft: per-module alpha / r with lora_alpha = r = most common rank (first seen wins ties)."""
    import collections

    r_modal = collections.Counter(ranks.values()).most_common(1)[0][0]
    return {m: r_modal / r for m, r in ranks.items()}


def lora_linear_runtime(x: torch.Tensor, weight: torch.Tensor, bias, adapters: Sequence[Tuple[torch.Tensor, torch.Tensor, float]]):
    """PEFT's runtime form for one Linear with several active adapters: adapters = [(A [r,K], B [N,r], scaling), ...]."""
    result = F.linear(x, weight, bias)
    for A, B, scaling in adapters:
        result = result + F.linear(F.linear(x, A), B) * scaling
    return result


def merged_weight(weight: torch.Tensor, adapters: Sequence[Tuple[torch.Tensor, torch.Tensor, float]]) -> torch.Tensor:
    """W + sum_i scaling_i * B_i A_i in the dtype of ``weight`` math promoted to fp32 (what a merge must equal)."""
    w = weight.float()
    for A, B, scaling in adapters:
        w = w + scaling * (B.float() @ A.float())
    return w


def merge_into_state_dict(w: Weights, lora: Weights, scale: float) -> Weights:
    """fp32 state dict with ``<module>.lora_A/B.weight`` of ``lora`` merged at ``scale`` -- exact-math equivalent of
    running every adapted Linear in the runtime form (linearity), used to drive oracle/wan_dit.py unchanged."""
    out = {k: v.clone().float() for k, v in w.items()}
    for k in lora:
        if k.endswith(".lora_A.weight"):
            m = k[:-len(".lora_A.weight")]
            out[m + ".weight"] = merged_weight(w[m + ".weight"], [(lora[k], lora[m + ".lora_B.weight"], scale)])
            if m + ".lora_B.bias" in lora:
                out[m + ".bias"] = w[m + ".bias"].float() + scale * lora[m + ".lora_B.bias"].float()
    return out


def make_lora(w: Weights, modules: Sequence[str], rank: int, seed: int = 0, std: float = 0.05) -> Weights:
    g = torch.Generator().manual_seed(seed)
    out: Weights = {}
    for m in modules:
        n, k = w[m + ".weight"].shape
        out[m + ".lora_A.weight"] = torch.randn(rank, k, generator=g) * std
        out[m + ".lora_B.weight"] = torch.randn(n, rank, generator=g) * std
    return out
