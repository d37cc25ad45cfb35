#!/usr/bin/env python
"""Secondary benchmark (BASELINE.json configs[2]): QwenImage-Edit-2509 1024x1024, 8-step Lightning: one DiT forward per step
(true CFG off with the Lightning LoRA merged) over 4096 noisy-latent tokens + 4096 tokens of one 1024x1024 reference image
(engine/qwenimage/edit_plus.py) + 512 text tokens, d = 3072, 24 heads, 60 dual-stream blocks, ONE B200.

    python scripts/bench_qwen.py [--steps K] [--warmup W] [--layers 60] [--ref-images 1]"""
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
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--layers", type=int, default=60)
    ap.add_argument("--ref-images", type=int, default=1)
    a = ap.parse_args()
    from apex_studio_b200 import ops
    from apex_studio_b200.qwenimage import QwenImageConfig, QwenImageTransformer2DModel

    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    m = QwenImageTransformer2DModel(QwenImageConfig(num_layers=a.layers)).init_random_weights(dev)
    g = torch.Generator(device=dev).manual_seed(42)
    shapes = [(1, 64, 64)] * (1 + a.ref_images)
    n_img, n_txt, d = 4096 * (1 + a.ref_images), 512, 3072
    x = torch.randn(1, n_img, 64, generator=g, device=dev).bfloat16()
    enc = torch.randn(1, n_txt, 3584, generator=g, device=dev).bfloat16()
    t = torch.tensor([0.5], device=dev)

    def fwd():
        return m(hidden_states=x, encoder_hidden_states=enc, timestep=t, img_shapes=[shapes], txt_seq_lens=[n_txt], return_dict=False)[0]

    for _ in range(a.warmup):
        y = fwd()
    ops.launch_count = 0
    from bench import ClockSampler   # nvidia-smi clocks / throttle reasons sampled during the timed region
    clk = ClockSampler(0)
    clk.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.steps):
        y = fwd()
    e1.record()
    torch.cuda.synchronize()
    clocks = clk.stop()
    ms = e0.elapsed_time(e1) / a.steps
    S = n_img + n_txt
    fl = a.layers * (24.0 * S * d * d + 4.0 * S * S * d)
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk)).get("bf16_tflops_sustained", 1376.1) if os.path.exists(pk) else 1376.1
    print(json.dumps({"metric": "denoise_steps_per_sec", "workload": "QwenImage-Edit-2509 1024x1024 (%d image + %d text tokens, "
                      "d=3072, 24 heads), %d dual-stream blocks, 1 forward per step" % (n_img, n_txt, a.layers),
                      "value": 1000.0 / ms, "unit": "steps/s", "ms_per_step": ms, "seconds_per_image_8_steps": 8 * ms / 1000.0,
                      "steps": a.steps, "warmup": a.warmup, "dtype": "bf16", "data": "synthetic",
                      "algorithmic_flops_per_step": fl, "tflops": fl / ms / 1e9, "frac_of_peak": fl / ms / 1e9 / peak, "peak": peak,
                      "gpu_launches_per_step": ops.launch_count // a.steps, "clocks": clocks, "finite": bool(torch.isfinite(y).all()),
                      "parameter_gb": m.parameter_bytes() / 1e9}))


if __name__ == "__main__":
    main()
