#!/usr/bin/env python
"""HunyuanVideo-1.5 VAE tiled decode of a 720p x 129-frame latent [32, 33, 45, 80] on ONE B200 (112 tiles of 8x8 latents,
all frames per tile), synthetic weights of the production widths.  `--tiles N` decodes only the first N tiles (timing
sample); the full decode is the default."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=0)
    a = ap.parse_args()
    from apex_studio_b200 import ops
    from apex_studio_b200.vae import AutoencoderKLHunyuanVideo15

    dev = torch.device("cuda:0")
    vae = AutoencoderKLHunyuanVideo15().init_random_weights(dev)
    vae.enable_tiling()
    z = torch.randn(1, 32, 33, 45, 80, device=dev, generator=torch.Generator(device=dev).manual_seed(1)).bfloat16()
    vae.decode_tile(z[0, :, :, :8, :8].contiguous())       # warm-up (one tile)
    torch.cuda.synchronize()
    ops.launch_count = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if a.tiles:
        grid = vae.tile_grid(45, 80)[: a.tiles]
        for i, j in grid:
            y = vae.decode_tile(z[0, :, :, i:i + 8, j:j + 8].contiguous())
        n_tiles = len(grid)
    else:
        y = vae.decode(z, return_dict=False)[0]
        n_tiles = len(vae.tile_grid(45, 80))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"metric": "vae_decode", "workload": "HunyuanVideo-1.5 VAE tiled decode 720p x 129f, %d of 112 tiles" % n_tiles,
                      "ms": ms, "ms_per_tile": ms / n_tiles, "extrapolated_full_decode_s": ms / n_tiles * 112 / 1000.0,
                      "frames": 129, "launches": ops.launch_count, "shape": list(y.shape), "finite": bool(torch.isfinite(y).all()),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}))


if __name__ == "__main__":
    main()
