#!/usr/bin/env python
"""Secondary benchmark (BASELINE.json configs[4]): HunyuanVideo-1.5 720p x 129 frames, one DiT forward (the engine runs
cond + uncond per step, engine/hunyuanvideo15/t2v.py:262-292) on ONE B200: latent grid 33 x 45 x 80 = 118,800 tokens + 1985
condition tokens (1000 MLLM + 256 ByT5 + 729 image), d = 2048, 16 heads, 54 dual-stream blocks (91 % attention).

    python scripts/bench_hy15.py [--steps K] [--warmup W] [--layers 54]
Prints one JSON line: forwards/s and denoise-steps/s (2 forwards), algorithmic TFLOP/s against the measured bf16 peak."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--layers", type=int, default=54)
    a = ap.parse_args()
    from apex_studio_b200 import ops
    from apex_studio_b200.hunyuanvideo15 import HunyuanVideo15Config, HunyuanVideo15Transformer3DModel

    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    m = HunyuanVideo15Transformer3DModel(HunyuanVideo15Config(num_layers=a.layers)).init_random_weights(dev)
    g = torch.Generator(device=dev).manual_seed(42)
    bf = torch.bfloat16
    x = torch.randn(1, 65, 33, 45, 80, generator=g, device=dev).to(bf)
    text = torch.randn(1, 1000, 3584, generator=g, device=dev).to(bf)
    text2 = torch.randn(1, 256, 1472, generator=g, device=dev).to(bf)
    img = torch.zeros(1, 729, 1152, device=dev, dtype=bf)      # t2v: all-zero image embeds (t2v.py:196-202)
    mask, mask2 = torch.ones(1, 1000), torch.ones(1, 256)
    mask[:, 300:], mask2[:, 64:] = 0, 0
    t = torch.tensor([750.0], device=dev, dtype=bf)

    def fwd():
        return m(x, t, text, mask, encoder_hidden_states_2=text2, encoder_attention_mask_2=mask2, image_embeds=img,
                 return_dict=False)[0]

    for _ in range(a.warmup):
        y = fwd()
    ops.launch_count = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.steps):
        y = fwd()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    S_lat, S_ctx, d = 118800, 1985, 2048
    S = S_lat + S_ctx
    fl = a.layers * (24.0 * S * d * d + 4.0 * S * S * d)
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk)).get("bf16_tflops_sustained", 1376.1) if os.path.exists(pk) else 1376.1
    print(json.dumps({"metric": "dit_forwards_per_sec", "workload": "HunyuanVideo-1.5 720p x 129f (118800 latent + 1985 condition "
                      "tokens, d=2048, 16 heads), %d dual-stream blocks" % a.layers, "value": 1000.0 / ms, "unit": "forwards/s",
                      "ms_per_forward": ms, "denoise_steps_per_sec_cfg": 500.0 / ms, "steps": a.steps, "warmup": a.warmup,
                      "dtype": "bf16", "data": "synthetic", "algorithmic_flops_per_forward": fl, "tflops": fl / ms / 1e9,
                      "frac_of_peak": fl / ms / 1e9 / peak, "peak": peak, "gpu_launches_per_forward": ops.launch_count // a.steps,
                      "finite": bool(torch.isfinite(y).all()), "parameter_gb": m.parameter_bytes() / 1e9}))


if __name__ == "__main__":
    main()
