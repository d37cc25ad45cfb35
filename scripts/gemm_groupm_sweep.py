#!/usr/bin/env python
"""GROUP_M (rasterisation) sweep of b200_linear on the Wan shapes, sustained timing; B200_LINEAR_GROUP_M is read per launch."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from apex_studio_b200 import ops
SHAPES = [("wan qkv", 75600, 15360, 5120, 0), ("wan out", 75600, 5120, 5120, 2), ("wan ff1", 75600, 13824, 5120, 1), ("wan ff2", 75600, 5120, 13824, 2)]
res = {}
for name, M, N, K, epi in SHAPES:
    x = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16(); g = torch.randn(N, device="cuda").bfloat16()
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    row = {}
    for gm in (4, 8, 12, 16, 24, 32, 48):
        os.environ["B200_LINEAR_GROUP_M"] = str(gm)
        f = lambda: ops.linear(x, w, b, epilogue=epi, out=out, gate=g if epi == 2 else None)
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(25): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 25
        row[gm] = round(2.0 * M * N * K / ms / 1e9, 1)
    ref = lambda: torch.matmul(x, w.t(), out=out)
    for _ in range(3): ref()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(25): ref()
    e1.record(); torch.cuda.synchronize()
    row["cublas"] = round(2.0 * M * N * K / (e0.elapsed_time(e1) / 25) / 1e9, 1)
    res[name] = row
    print(name, row, flush=True)
    del x, w, out
