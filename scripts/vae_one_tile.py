#!/usr/bin/env python
"""Decode ONE full-size Wan VAE tile (32 x 32 latents x 21 frames, base_dim 96) -- the unit the 28-tile decode repeats; the
command `ncu --metrics gpu__time_duration.sum` lists (per-kernel shares of the decode)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from apex_studio_b200 import ops
from apex_studio_b200.vae import AutoencoderKLWan
vae = AutoencoderKLWan().init_random_weights("cuda", seed=7)
z = torch.randn(16, 21, 32, 32, device="cuda").bfloat16()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
vae.decode_tile(z)
torch.cuda.synchronize()
n0 = ops.launch_count
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    out = vae.decode_tile(z)
e1.record()
torch.cuda.synchronize()
print(f"tile decode: {e0.elapsed_time(e1) / reps:.2f} ms, {(ops.launch_count - n0) // reps} launches, out {tuple(out.shape)}")
