#!/usr/bin/env python
"""One b200_linear shape, a few launches (ncu target): python scripts/gemm_one.py M N K [epi]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from apex_studio_b200 import ops
M, N, K = (int(x) for x in sys.argv[1:4])
epi = int(sys.argv[4]) if len(sys.argv) > 4 else 0
x = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
b = torch.randn(N, device="cuda").bfloat16()
out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.linear(x, w, b, epilogue=epi, out=out)
torch.cuda.synchronize()
print("done")
