"""Torch-tensor front end of the C ABI (include/apex_b200.h).

Each function extracts raw device pointers, element strides and the current CUDA stream from torch
tensors, calls the matching ``b200_*`` entry point of libapex_b200.so and maps error codes to the
exceptions the reference raises at the same place (``ValueError`` for shape/dtype/stride problems as in
apps/api/src/attention/functions.py:791-801, ``RuntimeError`` otherwise).  PyTorch is used for device
memory and streams only; all arithmetic happens in the CUDA kernels.  There is no fallback path.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib

EPI_BIAS, EPI_GELU_TANH, EPI_GATE_RES, EPI_BIAS_F32, EPI_SILU, EPI_GELU_ERF = 0, 1, 2, 3, 4, 5

#: number of kernels launched through this module since import / last reset (bench.py's gpu_launches)
launch_count = 0


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda_bf16(name: str, t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (the b200 path has no CPU fallback), got {t.device}")
    if t.dtype != torch.bfloat16:
        raise ValueError(f"{name} must be bfloat16, got {t.dtype}")


def _on_current_device(t: torch.Tensor) -> None:
    """Kernels are launched on the CURRENT device's current stream; a tensor that lives on another GPU would be
    dereferenced from the wrong context.  Fail loudly instead (callers use ``torch.cuda.device(t.device)``)."""
    if t.device.index != torch.cuda.current_device():
        raise ValueError(f"tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                         "wrap the call in torch.cuda.device(tensor.device)")


def _count(n: int = 1) -> None:
    global launch_count
    launch_count += n


# ---------------------------------------------------------------------------------------------------
# attention core
# ---------------------------------------------------------------------------------------------------
def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, softmax_scale: Optional[float] = None,
              out: Optional[torch.Tensor] = None, q_norm: Optional[tuple] = None) -> torch.Tensor:
    """softmax(q k^T * scale) v for q [B,H,Sq,128], k/v [B,H,Sk,128] (any strides with a contiguous last dim).

    Returns [B,H,Sq,128] as a transposed view of a [B,Sq,H,128] buffer -- the layout the caller wants next
    (reference: attention.py:397-401 immediately does ``.transpose(1, 2).flatten(2, 3)``).

    ``q_norm = (row_sumsq, n_parts, dim, eps)`` from :func:`linear_normw`: the RMS-norm of q over all heads' channels is
    applied to the logits row by row (``b200_attn_fwd_qnorm``) instead of to q in a separate pass.
    """
    for name, t in (("q", q), ("k", k), ("v", v)):
        _require_cuda_bf16(name, t)
        if t.dim() != 4:
            raise ValueError(f"{name} must be [B,H,S,D], got shape {tuple(t.shape)}")
    _on_current_device(q)
    B, H, Sq, D = q.shape
    Bk, Hk, Sk, Dk = k.shape
    if v.shape != k.shape or Bk != B or Hk != H or Dk != D:
        raise ValueError(f"shape mismatch: q {tuple(q.shape)}, k {tuple(k.shape)}, v {tuple(v.shape)}")
    if D != 128:
        raise ValueError(f"b200 attention supports head_dim 128 only, got {D}")

    def norm(t):
        # TMA needs a contiguous last dim and 16-byte aligned strides / base.
        if t.stride(3) != 1 or any(s % 8 for s in t.stride()[:3]) or t.data_ptr() % 16:
            return t.contiguous()
        return t

    q, k, v = norm(q), norm(k), norm(v)
    if out is None:
        out = torch.empty((B, Sq, H, D), dtype=q.dtype, device=q.device).transpose(1, 2)
    else:
        # the kernel stores 16-byte vectors along the head dim: a wrong-shaped / transposed `out` would corrupt memory
        _require_cuda_bf16("out", out)
        if out.device != q.device or tuple(out.shape) != (B, H, Sq, D) or out.stride(3) != 1:
            raise ValueError(f"out must be a bf16 [B,H,Sq,D] = {(B, H, Sq, D)} tensor on {q.device} with a contiguous last dim, "
                             f"got shape {tuple(out.shape)}, strides {out.stride()}, device {out.device}")
        if any(st % 8 for st in out.stride()[:3]) or out.data_ptr() % 16:
            raise ValueError("out needs 16-byte aligned strides and base address")
    scale = float(softmax_scale) if softmax_scale is not None else 1.0 / math.sqrt(D)
    lib = _lib.load()
    strides = (q.stride(0), q.stride(1), q.stride(2), k.stride(0), k.stride(1), k.stride(2),
               v.stride(0), v.stride(1), v.stride(2), out.stride(0), out.stride(1), out.stride(2))
    if q_norm is not None:
        sumsq, n_parts, dim, eps = q_norm
        if sumsq.dtype != torch.float32 or not sumsq.is_cuda or not sumsq.is_contiguous() or sumsq.numel() < B * Sq * n_parts:
            raise ValueError("q_norm row_sumsq must be a contiguous fp32 CUDA tensor of at least B * Sq * n_parts elements")
        rc = lib.b200_attn_fwd_qnorm(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, Sq, Sk, D, *strides, scale,
                                     sumsq.data_ptr(), int(n_parts), int(dim), float(eps), _stream())
        _lib.check(rc, "b200_attn_fwd_qnorm")
    else:
        rc = lib.b200_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, Sq, Sk, D, *strides, scale, _stream())
        _lib.check(rc, "b200_attn_fwd")
    _count()
    return out


def linear_normw(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], norm_weight: torch.Tensor, out: torch.Tensor,
                 row_sumsq: torch.Tensor) -> int:
    """out = bf16(q * norm_weight) with q = bf16(x @ weight^T + bias); ``row_sumsq`` (flat fp32, >= M * ceil(N / 64) elements)
    receives the sums of q^2 per row and column tile as [M, n_parts].  Returns n_parts, the number of parts per row the kernel
    wrote (pass it on as ``q_norm`` to :func:`attention`).  The RMS-norm of the Wan cross-attention query folded into its
    projection (attention.py:345-370)."""
    for name, t in (("x", x), ("weight", weight), ("norm_weight", norm_weight), ("out", out)):
        _require_cuda_bf16(name, t)
    if x.dim() != 2 or x.stride(1) != 1 or weight.stride(1) != 1 or out.dim() != 2 or out.stride(1) != 1:
        raise ValueError("x, weight and out must be 2-D with a contiguous last dim")
    M, K = x.shape
    N = weight.shape[0]
    if tuple(out.shape) != (M, N) or norm_weight.numel() != N or not norm_weight.is_contiguous():
        raise ValueError("shape mismatch")
    cap = row_sumsq.numel() // M
    if row_sumsq.dtype != torch.float32 or not row_sumsq.is_cuda or not row_sumsq.is_contiguous() or cap < (N + 63) // 64:
        raise ValueError("row_sumsq must be a contiguous fp32 CUDA tensor of at least M * ceil(N / 64) elements")
    import ctypes

    n_parts = ctypes.c_int(0)
    lib = _lib.load()
    rc = lib.b200_linear_normw(x.data_ptr(), weight.data_ptr(), _ptr(bias), norm_weight.data_ptr(), out.data_ptr(),
                               row_sumsq.data_ptr(), int(cap), ctypes.byref(n_parts), M, N, K, x.stride(0), weight.stride(0),
                               out.stride(0), _stream())
    _lib.check(rc, "b200_linear_normw")
    _count()
    return n_parts.value


# ---------------------------------------------------------------------------------------------------
# linear with fused epilogue
# ---------------------------------------------------------------------------------------------------
def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, *, epilogue: int = EPI_BIAS,
           out: Optional[torch.Tensor] = None, gate: Optional[torch.Tensor] = None,
           row_bias: bool = False) -> torch.Tensor:
    """out = epilogue(x @ weight^T + bias).  x [..., K] (last dim contiguous), weight [N, K].

    EPI_GATE_RES: ``out`` is the residual stream, updated in place: out += gate * (x @ W^T + bias).
    ``row_bias``: bias has one entry per ROW of x (transposed projections) instead of per output column.
    """
    _require_cuda_bf16("x", x)
    _require_cuda_bf16("weight", weight)
    _on_current_device(x)
    K = x.shape[-1]
    N = weight.shape[0]
    if weight.dim() != 2 or weight.shape[1] != K:
        raise ValueError(f"weight {tuple(weight.shape)} does not match x [..., {K}]")
    x2 = x.reshape(-1, K)
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    if weight.stride(1) != 1:
        weight = weight.contiguous()
    M = x2.shape[0]
    if epilogue == EPI_GATE_RES:
        if out is None:
            raise ValueError("EPI_GATE_RES updates `out` (the residual stream) in place; pass it")
        _require_cuda_bf16("out", out)
    if out is None:
        dt = torch.float32 if epilogue == EPI_BIAS_F32 else torch.bfloat16
        out = torch.empty(x.shape[:-1] + (N,), dtype=dt, device=x.device)
    o2 = out.view(-1, N) if out.is_contiguous() else out.reshape(-1, N)
    if o2.data_ptr() != out.data_ptr() or o2.stride(1) != 1:
        raise ValueError("`out` must be viewable as [M, N] with a contiguous last dim")
    if o2.shape[0] != M:
        raise ValueError(f"out rows {o2.shape[0]} != x rows {M}")
    if bias is not None:
        _require_cuda_bf16("bias", bias)
    if gate is not None:
        _require_cuda_bf16("gate", gate)
        gate = gate.reshape(-1)
        if gate.numel() != N or gate.stride(0) != 1:
            raise ValueError(f"gate must be a contiguous [{N}] vector")
    lib = _lib.load()
    rc = lib.b200_linear(x2.data_ptr(), weight.data_ptr(), _ptr(bias), o2.data_ptr(), _ptr(gate), M, N, K,
                         x2.stride(0), weight.stride(0), o2.stride(0), epilogue | (16 if row_bias else 0), _stream())
    _lib.check(rc, "b200_linear")
    _count()
    return out


# ---------------------------------------------------------------------------------------------------
# row kernels
# ---------------------------------------------------------------------------------------------------
def layernorm_modulate(x: torch.Tensor, scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None,
                       *, ln_weight: Optional[torch.Tensor] = None, ln_bias: Optional[torch.Tensor] = None,
                       eps: float = 1e-6, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """FP32LayerNorm(x) [* w + b] [* (1 + scale) + shift]; x [..., dim]; scale/shift [dim] or per-row [rows, dim]."""
    _require_cuda_bf16("x", x)
    dim = x.shape[-1]
    x2 = x.reshape(-1, dim)
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    rows = x2.shape[0]
    if out is None:
        out = torch.empty(x.shape, dtype=x.dtype, device=x.device)
    o2 = out.view(-1, dim)
    mod_stride = 0
    if scale is not None:
        _require_cuda_bf16("scale", scale)
        _require_cuda_bf16("shift", shift)
        if scale.numel() == dim:
            scale, shift = scale.reshape(dim).contiguous(), shift.reshape(dim).contiguous()
        elif scale.numel() == rows * dim:
            scale, shift = scale.reshape(rows, dim).contiguous(), shift.reshape(rows, dim).contiguous()
            mod_stride = dim
        else:
            raise ValueError(f"scale/shift must have {dim} or {rows * dim} elements, got {scale.numel()}")
    for name, t in (("ln_weight", ln_weight), ("ln_bias", ln_bias)):
        if t is not None:
            _require_cuda_bf16(name, t)
    lib = _lib.load()
    rc = lib.b200_layernorm_modulate(x2.data_ptr(), o2.data_ptr(), _ptr(scale), _ptr(shift), _ptr(ln_weight),
                                     _ptr(ln_bias), rows, dim, x2.stride(0), o2.stride(0), mod_stride, float(eps),
                                     _stream())
    _lib.check(rc, "b200_layernorm_modulate")
    _count()
    return out


def rmsnorm_rope_(x: torch.Tensor, weight: Optional[torch.Tensor], rope: Optional[torch.Tensor], heads: int,
                  eps: float = 1e-6, *, norm: bool = True) -> torch.Tensor:
    """In place: x <- RoPE(RMSNorm_over_all_channels(x) * weight).  x [rows, heads*head_dim] (row stride free).

    rope: bf16 [rows, head_dim] interleaved (cos0, sin0, cos1, sin1, ...) or None.
    """
    _require_cuda_bf16("x", x)
    if x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("x must be [rows, dim] with a contiguous last dim")
    rows, dim = x.shape
    if dim % heads:
        raise ValueError(f"dim {dim} not divisible by heads {heads}")
    head_dim = dim // heads
    if weight is not None:
        _require_cuda_bf16("weight", weight)
    if rope is not None:
        _require_cuda_bf16("rope", rope)
        if tuple(rope.shape) != (rows, head_dim) or not rope.is_contiguous():
            raise ValueError(f"rope must be contiguous [{rows}, {head_dim}], got {tuple(rope.shape)}")
    lib = _lib.load()
    rc = lib.b200_rmsnorm_rope(x.data_ptr(), _ptr(weight), _ptr(rope), rows, heads, head_dim, x.stride(0),
                               float(eps) if norm else -1.0, _stream())
    _lib.check(rc, "b200_rmsnorm_rope")
    _count()
    return x


def rmsnorm_rope_batched_(x: torch.Tensor, weight: torch.Tensor, rope: Optional[torch.Tensor], heads: int, eps: float,
                          n_batch: int, x_batch_stride: int) -> torch.Tensor:
    """rmsnorm_rope_ over ``n_batch`` column blocks of the same rows in ONE launch: block b = x[:, b * x_batch_stride :
    b * x_batch_stride + dim] normalised with weight[b] (weight [n_batch, dim] contiguous).  ``x`` is the FIRST block
    (a [rows, dim] view); the caller guarantees the other blocks exist behind it in the same rows."""
    _require_cuda_bf16("x", x)
    _require_cuda_bf16("weight", weight)
    if x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("x must be [rows, dim] with a contiguous last dim")
    rows, dim = x.shape
    if dim % heads or tuple(weight.shape) != (n_batch, dim) or not weight.is_contiguous():
        raise ValueError(f"weight must be a contiguous [{n_batch}, {dim}] tensor and dim divisible by heads")
    if (n_batch - 1) * x_batch_stride + dim > x.stride(0):
        raise ValueError("the column blocks do not fit into one row of the underlying buffer")
    head_dim = dim // heads
    if rope is not None:
        _require_cuda_bf16("rope", rope)
        if tuple(rope.shape) != (rows, head_dim) or not rope.is_contiguous():
            raise ValueError(f"rope must be contiguous [{rows}, {head_dim}], got {tuple(rope.shape)}")
    lib = _lib.load()
    rc = lib.b200_rmsnorm_rope_batched(x.data_ptr(), weight.data_ptr(), _ptr(rope), rows, heads, head_dim, x.stride(0), float(eps),
                                       n_batch, x_batch_stride, dim, _stream())
    _lib.check(rc, "b200_rmsnorm_rope_batched")
    _count()
    return x


def rmsnorm_rope_scatter(x: torch.Tensor, weight: Optional[torch.Tensor], rope: Optional[torch.Tensor], heads: int,
                         eps: float, peers, n_peers: int, dst_elem_offset: int, row0: int, *, norm: bool = True) -> None:
    """rmsnorm_rope_ whose result is stored into the peer GPUs' receive planes (tokens -> heads exchange fused into
    the kernel); ``peers`` is a ctypes array of ``n_peers`` device pointers (parallel.PeerExchange)."""
    _require_cuda_bf16("x", x)
    if x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("x must be [rows, dim] with a contiguous last dim")
    rows, dim = x.shape
    lib = _lib.load()
    rc = lib.b200_rmsnorm_rope_scatter(x.data_ptr(), _ptr(weight), _ptr(rope), rows, heads, dim // heads, x.stride(0),
                                       float(eps) if norm else -1.0, peers, n_peers, dst_elem_offset, row0, _stream())
    _lib.check(rc, "b200_rmsnorm_rope_scatter")
    _count()


def attention_scatter(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o_peers, n_peers: int, rows_per_rank: int,
                      head_off: int, o_row_stride: int, softmax_scale: Optional[float] = None, *, rep_rows: int = 0,
                      rep_first: bool = False) -> None:
    """attention for q/k/v [1,H,S,128] whose output rows are stored into the token-owning peers' [S/P, H_total*128]
    buffers (heads -> tokens exchange fused into the attention epilogue).  Joint sequences of the dual-stream families:
    ``rep_rows`` rows (the replicated text stream; first if ``rep_first`` else last) are stored to every peer, whose buffer
    is [rep_rows + S_img/P, H_total*128] in its local joint order."""
    for name, t in (("q", q), ("k", k), ("v", v)):
        _require_cuda_bf16(name, t)
        if t.dim() != 4 or t.stride(3) != 1 or any(st % 8 for st in t.stride()[:3]) or t.data_ptr() % 16:
            raise ValueError(f"{name} must be [1,H,S,128] with a contiguous last dim and 16-byte aligned strides / base")
    _on_current_device(q)
    B, H, Sq, D = q.shape
    Sk = k.shape[2]
    if B != 1 or D != 128:
        raise ValueError("attention_scatter needs batch 1 and head_dim 128")
    if tuple(k.shape) != (1, H, Sk, D) or v.shape != k.shape:
        raise ValueError(f"shape mismatch: q {tuple(q.shape)}, k {tuple(k.shape)}, v {tuple(v.shape)}")
    scale = float(softmax_scale) if softmax_scale is not None else 1.0 / math.sqrt(D)
    lib = _lib.load()
    rc = lib.b200_attn_fwd_scatter_joint(q.data_ptr(), k.data_ptr(), v.data_ptr(), H, Sq, Sk, D, q.stride(1), q.stride(2),
                                         k.stride(1), k.stride(2), v.stride(1), v.stride(2), o_peers, n_peers, rows_per_rank,
                                         head_off, D, o_row_stride, int(rep_rows), 1 if rep_first else 0, scale, _stream())
    _lib.check(rc, "b200_attn_fwd_scatter_joint")
    _count()


def gate_residual_(h: torch.Tensor, y: torch.Tensor, gate: Optional[torch.Tensor] = None) -> torch.Tensor:
    """In place: h += y * gate (gate [dim] or None)."""
    _require_cuda_bf16("h", h)
    _require_cuda_bf16("y", y)
    dim = h.shape[-1]
    h2, y2 = h.view(-1, dim), y.reshape(-1, dim)
    if gate is not None:
        _require_cuda_bf16("gate", gate)
        gate = gate.reshape(dim).contiguous()
    lib = _lib.load()
    rc = lib.b200_gate_residual(h2.data_ptr(), y2.data_ptr(), _ptr(gate), h2.shape[0], dim, h2.stride(0),
                                y2.stride(0), _stream())
    _lib.check(rc, "b200_gate_residual")
    _count()
    return h


def cfg_combine(cond: torch.Tensor, uncond: torch.Tensor, guidance_scale: float) -> torch.Tensor:
    """uncond + g * (cond - uncond) with the reference's bf16 rounding, bf16 out (wan/shared/__init__.py:565)."""
    _require_cuda_bf16("cond", cond)
    _require_cuda_bf16("uncond", uncond)
    cond, uncond = cond.contiguous(), uncond.contiguous()
    out = torch.empty(cond.shape, dtype=torch.bfloat16, device=cond.device)
    lib = _lib.load()
    rc = lib.b200_cfg_combine(cond.data_ptr(), uncond.data_ptr(), out.data_ptr(), float(guidance_scale),
                              cond.numel(), _stream())
    _lib.check(rc, "b200_cfg_combine")
    _count()
    return out


# ---------------------------------------------------------------------------------------------------
# dual-stream (MMDiT) families: Flux, HunyuanVideo-1.5, QwenImage
# ---------------------------------------------------------------------------------------------------
NORM_NONE, NORM_TORCH_RMS, NORM_INPLACE_RMS, NORM_DIFFUSERS_RMS = 0, 1, 2, 3


def adaln_zero_modulate(x: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, *, eps: float = 1e-6,
                        out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``LayerNorm(x) * (1 + scale) + shift`` with the rounding points of diffusers' AdaLayerNormZero family
    (flux/base/model.py:266-272,297-300); x [rows, dim] (row stride free), scale/shift [dim]."""
    _require_cuda_bf16("x", x)
    _require_cuda_bf16("scale", scale)
    _require_cuda_bf16("shift", shift)
    if x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("x must be [rows, dim] with a contiguous last dim")
    rows, dim = x.shape
    scale, shift = scale.reshape(-1), shift.reshape(-1)
    if scale.numel() != dim or shift.numel() != dim or scale.stride(0) != 1 or shift.stride(0) != 1:
        raise ValueError(f"scale/shift must be contiguous [{dim}] vectors")
    if out is None:
        out = torch.empty((rows, dim), dtype=x.dtype, device=x.device)
    if tuple(out.shape) != (rows, dim) or out.stride(1) != 1:
        raise ValueError(f"out must be [{rows}, {dim}] with a contiguous last dim")
    _require_cuda_bf16("out", out)
    rc = _lib.load().b200_adaln_zero_modulate(x.data_ptr(), out.data_ptr(), scale.data_ptr(), shift.data_ptr(), rows, dim,
                                              x.stride(0), out.stride(0), float(eps), _stream())
    _lib.check(rc, "b200_adaln_zero_modulate")
    _count()
    return out


def headnorm_rope_(q: torch.Tensor, k: Optional[torch.Tensor], wq: Optional[torch.Tensor], wk: Optional[torch.Tensor],
                   rope: Optional[torch.Tensor], heads: int, eps: float = 1e-6, norm_mode: int = NORM_TORCH_RMS) -> None:
    """In place on q (and k): per-head RMS-norm over head_dim = 128, then rotation of (even, odd) channel pairs.

    q, k: [rows, heads*128] column blocks of one buffer (same row stride); rope: fp32 [rows, 64, 2] = (cos, sin)."""
    _require_cuda_bf16("q", q)
    if q.dim() != 2 or q.stride(1) != 1 or q.shape[1] != heads * 128:
        raise ValueError(f"q must be [rows, {heads * 128}] with a contiguous last dim (head_dim 128 only)")
    rows = q.shape[0]
    if k is not None:
        _require_cuda_bf16("k", k)
        if tuple(k.shape) != tuple(q.shape) or k.stride() != q.stride():
            raise ValueError("k must have q's shape and strides")
    for name, t in (("wq", wq), ("wk", wk)):
        if t is not None:
            _require_cuda_bf16(name, t)
            if t.numel() != 128 or not t.is_contiguous():
                raise ValueError(f"{name} must be a contiguous [128] gain")
    if rope is not None:
        if not rope.is_cuda or rope.dtype != torch.float32 or tuple(rope.shape) != (rows, 64, 2) or not rope.is_contiguous():
            raise ValueError(f"rope must be a contiguous CUDA float32 [{rows}, 64, 2] (cos, sin) table")
    if norm_mode not in (0, 1, 2, 3):
        raise ValueError(f"norm_mode {norm_mode}")
    rc = _lib.load().b200_headnorm_rope(q.data_ptr(), _ptr(k), _ptr(wq), _ptr(wk), _ptr(rope), rows, heads, 128,
                                        q.stride(0), float(eps), norm_mode, _stream())
    _lib.check(rc, "b200_headnorm_rope")
    _count()


def rmsnorm_rows(x: torch.Tensor, weight: Optional[torch.Tensor], eps: float = 1e-6, norm_mode: int = NORM_DIFFUSERS_RMS,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """RMS-norm with gain over the last dim of x [rows, dim] (diffusers RMSNorm of the QwenImage text stream,
    qwenimage/base/model.py:826,920); rounding points selected by ``norm_mode`` (see headnorm_rope_)."""
    _require_cuda_bf16("x", x)
    if x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("x must be [rows, dim] with a contiguous last dim")
    rows, dim = x.shape
    if weight is not None:
        _require_cuda_bf16("weight", weight)
        if weight.numel() != dim or not weight.is_contiguous():
            raise ValueError(f"weight must be a contiguous [{dim}] gain")
    if out is None:
        out = torch.empty((rows, dim), dtype=x.dtype, device=x.device)
    _require_cuda_bf16("out", out)
    if tuple(out.shape) != (rows, dim) or out.stride(1) != 1:
        raise ValueError(f"out must be [{rows}, {dim}] with a contiguous last dim")
    rc = _lib.load().b200_rmsnorm_rows(x.data_ptr(), out.data_ptr(), _ptr(weight), rows, dim, x.stride(0), out.stride(0),
                                       float(eps), norm_mode, _stream())
    _lib.check(rc, "b200_rmsnorm_rows")
    _count()
    return out


def swiglu(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """silu(x[:, :inner]) * x[:, inner:] (Flux2SwiGLU, flux2/base/model.py:91-105); x [rows, 2*inner]."""
    _require_cuda_bf16("x", x)
    if x.dim() != 2 or x.stride(1) != 1 or x.shape[1] % 2:
        raise ValueError("x must be [rows, 2*inner] with a contiguous last dim")
    rows, inner = x.shape[0], x.shape[1] // 2
    if out is None:
        out = torch.empty((rows, inner), dtype=x.dtype, device=x.device)
    _require_cuda_bf16("out", out)
    if tuple(out.shape) != (rows, inner) or out.stride(1) != 1:
        raise ValueError(f"out must be [{rows}, {inner}] with a contiguous last dim")
    rc = _lib.load().b200_swiglu(x.data_ptr(), out.data_ptr(), rows, inner, x.stride(0), out.stride(0), _stream())
    _lib.check(rc, "b200_swiglu")
    _count()
    return out


# ---------------------------------------------------------------------------------------------------
# composite call sites (SURVEY.md section 8b granularity): one C call enqueues the kernels of one reference call site
# ---------------------------------------------------------------------------------------------------
def qkv_rmsnorm_rope(x: torch.Tensor, w_qkv: torch.Tensor, b_qkv: Optional[torch.Tensor], wq_norm: torch.Tensor,
                     wk_norm: torch.Tensor, rope: Optional[torch.Tensor], heads: int, eps: float,
                     out: torch.Tensor) -> torch.Tensor:
    """attention.py:345-370 in one call: out[rows, 3*dim] = x @ W_qkv^T + b, then RMS-norm-across-heads + RoPE in place
    on the q and k column blocks.  Identical launches (and bits) to linear + 2 x rmsnorm_rope_."""
    for name, t in (("x", x), ("w_qkv", w_qkv), ("wq_norm", wq_norm), ("wk_norm", wk_norm), ("out", out)):
        _require_cuda_bf16(name, t)
    rows, dim = x.shape
    if x.stride(1) != 1 or tuple(w_qkv.shape) != (3 * dim, dim) or not w_qkv.is_contiguous():
        raise ValueError("x must be [rows, dim] (contiguous rows), w_qkv a contiguous [3*dim, dim]")
    if tuple(out.shape) != (rows, 3 * dim) or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous [{rows}, {3 * dim}] buffer")
    if rope is not None and (tuple(rope.shape) != (rows, dim // heads) or not rope.is_contiguous()):
        raise ValueError(f"rope must be contiguous [{rows}, {dim // heads}]")
    lib = _lib.load()
    rc = lib.b200_qkv_rmsnorm_rope(x.data_ptr(), w_qkv.data_ptr(), _ptr(b_qkv), wq_norm.data_ptr(), wk_norm.data_ptr(),
                                   _ptr(rope), out.data_ptr(), rows, dim, heads, x.stride(0), float(eps), _stream())
    _lib.check(rc, "b200_qkv_rmsnorm_rope")
    stacked = wk_norm.data_ptr() == wq_norm.data_ptr() + 2 * dim      # [2, dim] norm weights -> q and k share one launch
    _count(2 if stacked else 3)
    return out


def mlp_gelu_(h: torch.Tensor, x: torch.Tensor, w1: torch.Tensor, b1: Optional[torch.Tensor], w2: torch.Tensor,
              b2: Optional[torch.Tensor], gate: Optional[torch.Tensor], workspace: torch.Tensor) -> torch.Tensor:
    """model.py:1265-1279 in one call, in place on the residual stream: h += gate * (W2 gelu_tanh(W1 x + b1) + b2)."""
    for name, t in (("h", h), ("x", x), ("w1", w1), ("w2", w2), ("workspace", workspace)):
        _require_cuda_bf16(name, t)
    rows, dim = x.shape
    ffn = w1.shape[0]
    if not (x.is_contiguous() and h.is_contiguous() and w1.is_contiguous() and w2.is_contiguous()):
        raise ValueError("mlp_gelu_ needs contiguous tensors")
    if tuple(h.shape) != (rows, dim) or tuple(w1.shape) != (ffn, dim) or tuple(w2.shape) != (dim, ffn):
        raise ValueError("shape mismatch between h, x, w1, w2")
    if workspace.numel() < rows * ffn or not workspace.is_contiguous():
        raise ValueError(f"workspace must hold [{rows}, {ffn}] bf16")
    if gate is not None:
        _require_cuda_bf16("gate", gate)
        gate = gate.reshape(dim).contiguous()
    lib = _lib.load()
    rc = lib.b200_mlp_gelu(x.data_ptr(), w1.data_ptr(), _ptr(b1), w2.data_ptr(), _ptr(b2), _ptr(gate), h.data_ptr(),
                           workspace.data_ptr(), rows, dim, ffn, _stream())
    _lib.check(rc, "b200_mlp_gelu")
    _count(2)
    return h
