#!/usr/bin/env python
"""Sustained timing (about 1 s windows) of the Wan GEMMs with their production epilogues: to_out / ffn.2 with the gated
residual update (EPI_GATE_RES, read-modify-write of h), ffn.0 with GELU, QKV plain.  APEX_B200_LIB selects the library."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from apex_studio_b200 import ops
M = 75600
res = {"lib": os.path.basename(os.environ.get("APEX_B200_LIB", "libapex_b200.so"))}
for name, N, K, epi in (("to_out_gate_res", 5120, 5120, ops.EPI_GATE_RES), ("ffn2_gate_res", 5120, 13824, ops.EPI_GATE_RES),
                        ("ffn0_gelu", 13824, 5120, ops.EPI_GELU_TANH), ("qkv", 15360, 5120, 0)):
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16()
    g = torch.randn(N, device="cuda").bfloat16() * 0.01
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    kw = dict(epilogue=epi, out=out)
    if epi == ops.EPI_GATE_RES:
        kw["gate"] = g
    for _ in range(3):
        ops.linear(x, w, b, **kw)
    torch.cuda.synchronize()
    flop = 2.0 * M * N * K
    iters = max(10, int(1.0 / (flop / 1.2e15)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.linear(x, w, b, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    res[name] = {"ms": round(ms, 3), "tflops": round(flop / ms / 1e9, 1)}
    del x, w, out
print(json.dumps(res))
