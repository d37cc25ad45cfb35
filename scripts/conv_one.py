#!/usr/bin/env python
"""One 3x3x3 causal conv launch at a Wan VAE stage shape (default: the full-resolution 96 -> 96 stage of a 32x32x21 tile:
81 x 256 x 256 pixels) after warm-up launches -- the command ncu captures; prints the CUDA-event time and TFLOP/s."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from apex_studio_b200.vae.wan import conv3d_cl
T, H, W, cin, cout = (int(a) for a in sys.argv[1:6]) if len(sys.argv) > 5 else (81, 256, 256, 96, 96)
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 5
x = torch.randn(T, H, W, cin, device="cuda").bfloat16()
w = (torch.randn(27 * cout, cin, device="cuda") * (27 * cin) ** -0.5).bfloat16()
b = torch.randn(cout, device="cuda").bfloat16()
res = torch.randn(T, H, W, cout, device="cuda").bfloat16()
out = torch.empty(T, H, W, cout, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    conv3d_cl(x, w, b, (3, 3, 3), cout, residual=res, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    conv3d_cl(x, w, b, (3, 3, 3), cout, residual=res, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"conv {T}x{H}x{W} {cin}->{cout}: {ms:.3f} ms, {2.0 * T * H * W * 27 * cin * cout / ms / 1e9:.0f} TFLOP/s")
