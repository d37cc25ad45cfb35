"""The attention operator boundary (boundary #1).

``attention_register`` mirrors apps/api/src/attention/functions.py:84; the ``"b200"`` backend has the
signature every reference backend has (functions.py:338-377)::

    fn(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False, softmax_scale=None, **kwargs) -> Tensor

q [B,H,Sq,D], k/v [B,H,Sk,D] with arbitrary strides (callers pass ``.transpose(1, 2)`` views); returns
[B,H,Sq,D] in q's dtype (a transposed view, as the reference's ``flash`` backend returns, functions.py:838).
It tolerates the extra kwargs the reference's verifier and smoke test pass (``cu_seqlens_q/k``,
``max_seqlen_q/k``, ``default_dtype``, ``attention_mask``; functions.py:2075-2083).  There is no fallback:
a mask, dropout, causal flag or unsupported head_dim raises instead of degrading to ``sdpa``.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .register import FunctionRegister

attention_register = FunctionRegister()


def _b200_available() -> bool:
    try:
        from . import _lib

        _lib.load()
    except Exception:
        return False
    return torch.cuda.is_available() and torch.cuda.get_device_capability()[0] == 10


def b200_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, attn_mask: Optional[torch.Tensor] = None,
                   dropout_p: float = 0.0, is_causal: bool = False, softmax_scale: Optional[float] = None,
                   **kwargs) -> torch.Tensor:
    if attn_mask is not None or kwargs.get("attention_mask") is not None:
        raise ValueError("b200 attention: masks are not supported (the hot paths pass attn_mask=None)")
    if dropout_p:
        raise ValueError("b200 attention: dropout is not supported (inference only)")
    if is_causal:
        raise ValueError("b200 attention: is_causal=True is not supported")
    if q.dtype not in (torch.bfloat16,):
        raise ValueError(f"b200 attention: bfloat16 only, got {q.dtype}")
    return ops.attention(q, k, v, softmax_scale=softmax_scale)


def register_b200(registry=None, *, name: str = "b200", make_default: bool = False):
    """Register the backend into ``registry`` (the reference's own ``attention_register`` when running inside
    the server, this module's mirror otherwise)."""
    reg = attention_register if registry is None else registry
    reg(name, overwrite=True, available=_b200_available())(b200_attention)
    if make_default:
        reg.set_default(name)
    return reg


register_b200()
