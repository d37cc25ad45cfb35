"""Wan 3-axis rotary table, built the way the reference builds it (transformer/wan/base/model.py:826-951):
float64 angles pos * theta^(-2i/d) for the time / height / width slices of the head dim
(h = w = 2*(head_dim//6), t = rest), time rows taken from index 1 of a table that starts at t = -1,
concatenated per token in (frame, row, col) order.  The kernel wants the values the reference multiplies
with -- cos/sin ALREADY cast to bf16 (transformer/efficiency/ops.py:157-158) -- interleaved as
(cos0, sin0, cos1, sin1, ...) so that one 16-byte load covers four (re, im) pairs.
"""
from __future__ import annotations

from typing import Tuple

import torch


def _angles(dim: int, length: int, theta: float, start: int) -> torch.Tensor:
    inv = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float64) / dim))
    pos = torch.arange(start, start + length, dtype=torch.float64)
    return torch.outer(pos, inv)


def wan_rope_angles(head_dim: int, grid: Tuple[int, int, int], max_seq_len: int = 1024, theta: float = 10000.0,
                    time_offset: int = -1) -> torch.Tensor:
    """float64 [F*H*W, head_dim/2] rotation angles."""
    f, h, w = grid
    hw_dim = 2 * (head_dim // 6)
    t_dim = head_dim - 2 * hw_dim
    t_rows = _angles(t_dim, max_seq_len + (1 if time_offset < 0 else 0), theta, time_offset)
    t0 = 1 if time_offset < 0 else 0
    if t0 + f > t_rows.shape[0] or h > max_seq_len or w > max_seq_len:
        raise IndexError(f"rope grid {grid} exceeds max_seq_len {max_seq_len}")
    ta = t_rows[t0:t0 + f].view(f, 1, 1, -1).expand(f, h, w, -1)
    ha = _angles(hw_dim, max_seq_len, theta, 0)[:h].view(1, h, 1, -1).expand(f, h, w, -1)
    wa = _angles(hw_dim, max_seq_len, theta, 0)[:w].view(1, 1, w, -1).expand(f, h, w, -1)
    return torch.cat([ta, ha, wa], dim=-1).reshape(f * h * w, head_dim // 2)


def wan_rope_table_bf16(head_dim: int, grid: Tuple[int, int, int], device, max_seq_len: int = 1024,
                        theta: float = 10000.0) -> torch.Tensor:
    """bf16 [F*H*W, head_dim] = (cos0, sin0, cos1, sin1, ...); float64 -> bf16 in ONE rounding like the reference."""
    ang = wan_rope_angles(head_dim, grid, max_seq_len, theta)
    # torch.polar(1, ang) in the reference == (cos, sin) in float64
    table = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).reshape(ang.shape[0], head_dim)
    return table.to(torch.bfloat16).to(device).contiguous()
