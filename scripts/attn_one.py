#!/usr/bin/env python
"""One Wan self-attention launch (heads x S^2) after a warm-up launch -- the command ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from apex_studio_b200 import ops
H = int(sys.argv[1]) if len(sys.argv) > 1 else 40
S = int(sys.argv[2]) if len(sys.argv) > 2 else 75600
Sk = int(sys.argv[3]) if len(sys.argv) > 3 else S
q = torch.randn(1, H, S, 128, device="cuda", dtype=torch.bfloat16)
k, v = (torch.randn(1, H, Sk, 128, device="cuda", dtype=torch.bfloat16) for _ in range(2))
o = torch.empty_like(q)
for _ in range(3):
    ops.attention(q, k, v, out=o)
torch.cuda.synchronize()
print("done")
