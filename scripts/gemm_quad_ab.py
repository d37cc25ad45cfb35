#!/usr/bin/env python
"""A/B of the GEMM cluster forms (B200_LINEAR_QUAD read once per process): correctness vs fp32 math on a few shapes, then long-window
sustained throughput on the Wan shapes next to cuBLAS."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from apex_studio_b200 import ops
res = {"quad": os.environ.get("B200_LINEAR_QUAD", "default"), "tall": os.environ.get("B200_LINEAR_TALL", "default"), "tma_store": os.environ.get("B200_LINEAR_TMA_STORE", "default")}
for (M, N, K, epi) in [(300, 520, 264, 0), (1000, 5120, 512, 1), (2048, 768, 512, 2), (4096, 1024, 1024, 0), (777, 1304, 320, 2), (75600, 64, 640, 1), (333, 200, 136, 0), (1000, 328, 64, 4)]:
    torch.manual_seed(M)
    x = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16(); b = torch.randn(N, device="cuda").bfloat16()
    acc = x.float() @ w.float().t() + b.float()
    if epi == 0: ref, out = acc, ops.linear(x, w, b)
    elif epi == 4: ref, out = torch.nn.functional.silu(acc), ops.linear(x, w, b, epilogue=4)
    elif epi == 1: ref, out = torch.nn.functional.gelu(acc, approximate="tanh"), ops.linear(x, w, b, epilogue=ops.EPI_GELU_TANH)
    else:
        h = torch.randn(M, N, device="cuda").bfloat16(); g = torch.randn(N, device="cuda").bfloat16()
        ref, out = h.float() + g.float() * acc, h.clone(); ops.linear(x, w, b, epilogue=ops.EPI_GATE_RES, out=out, gate=g)
    res[f"rel_{M}x{N}x{K}_{epi}"] = round(((out.float() - ref).norm() / ref.norm()).item(), 5)
print(json.dumps(res), flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "bench":
    def timed(f, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): f()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    for name, M, N, K, epi in [("wan qkv", 75600, 15360, 5120, 0), ("wan ff1", 75600, 13824, 5120, 1), ("wan q2", 75600, 5120, 5120, 0)]:
        x = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
        b = torch.randn(N, device="cuda").bfloat16(); g = torch.randn(N, device="cuda").bfloat16(); out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        f = lambda: ops.linear(x, w, b, epilogue=epi, out=out, gate=g if epi == 2 else None)
        fc = lambda: torch.matmul(x, w.t(), out=out)
        row = {}
        for nm, fn in (("ours", f), ("cublas", fc), ("ours2", f)):
            t0 = time.time()
            while time.time() - t0 < 0.5: timed(fn, 4)
            ms, t0 = [], time.time()
            while time.time() - t0 < 1.5: ms.append(timed(fn, 4))
            row[nm] = round(2.0 * M * N * K / (sum(ms) / len(ms)) / 1e9, 1)
        print(name, json.dumps(row), flush=True)
        del x, w, out
