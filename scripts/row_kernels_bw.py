#!/usr/bin/env python
"""Achieved HBM bandwidth of the HBM-bound row kernels added for the dual-stream families and the HunyuanVideo-1.5 VAE
(CUDA events, back-to-back launches on buffers larger than L2): algorithmic bytes (each element read once, written once)
divided by the launch time, against the measured copy bandwidth in MEASURED_PEAKS.json.  With B200_PROFILE=1 exactly one
launch of each kernel runs between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from apex_studio_b200 import ops  # noqa: E402
from apex_studio_b200.vae.hunyuanvideo15 import dcae_upsample_cl, pad_norm_silu_cl  # noqa: E402

dev = "cuda"
bf = torch.bfloat16
prof = os.environ.get("B200_PROFILE") == "1"


def timeit(fn, iters=20):
    if prof:
        fn()
        return float("nan")
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk)).get("hbm_gbs", 6552.6) if os.path.exists(pk) else 6552.6
    rows = []
    if prof:
        torch.cuda.cudart().cudaProfilerStart()
    # per-head RMS-norm + RoPE on q and k of the HunyuanVideo-1.5 latent stream: [118800, 3*2048] buffer, q and k blocks
    S, d, H = 118800, 2048, 16
    qkv = torch.randn(S, 3 * d, device=dev).to(bf)
    wq = torch.ones(128, device=dev, dtype=bf)
    rope = torch.randn(S, 64, 2, device=dev)
    ms = timeit(lambda: ops.headnorm_rope_(qkv[:, :d], qkv[:, d:2 * d], wq, wq, rope, H, 1e-6, ops.NORM_INPLACE_RMS))
    rows.append(("headnorm_rope_kernel (HY-1.5 q,k: 118800 x 2 x 2048)", 2 * S * d * 2 * 2 + S * 512, ms))
    del qkv, rope
    # AdaLayerNormZero modulate on the FLUX joint stream x 16 (larger than L2): [73728, 3072]
    x = torch.randn(73728, 3072, device=dev).to(bf)
    y = torch.empty_like(x)
    sc = torch.randn(3072, device=dev).to(bf)
    ms = timeit(lambda: ops.adaln_zero_modulate(x, sc, sc, out=y))
    rows.append(("layernorm_modulate_kernel<.,1> (adaLN-zero, 73728 x 3072)", 2 * x.numel() * 2, ms))
    ms = timeit(lambda: ops.rmsnorm_rows(x, sc, 1e-6, ops.NORM_DIFFUSERS_RMS, out=y))
    rows.append(("rmsnorm_rows_kernel (73728 x 3072)", 2 * x.numel() * 2, ms))
    x2 = torch.randn(73728, 2 * 3072, device=dev).to(bf)
    ms = timeit(lambda: ops.swiglu(x2, out=y))
    rows.append(("swiglu_kernel (73728 x 2 x 3072)", 3 * x.numel() * 2, ms))
    del x, y, x2
    # HunyuanVideo-1.5 VAE last stage: [129, 128, 128, 128] -> padded [131, 130, 130, 128] with RMS-norm + SiLU
    a = torch.randn(129, 128, 128, 128, device=dev).to(bf)
    g = torch.ones(128, device=dev, dtype=bf)
    ms = timeit(lambda: pad_norm_silu_cl(a, g, True))
    rows.append(("pad_norm_silu_kernel (129x128x128x128 -> 131x130x130x128)", (a.numel() + 131 * 130 * 130 * 128) * 2, ms))
    # DCAE upsample 256 -> 128 channels at [129, 64, 64]: h [129,64,64,512], x [129,64,64,256] -> [129,128,128,128]
    h = torch.randn(129, 64, 64, 512, device=dev).to(bf)
    xs = torch.randn(129, 64, 64, 256, device=dev).to(bf)
    ms = timeit(lambda: dcae_upsample_cl(h, xs, 128, False))
    rows.append(("dcae_upsample_kernel (129x64x64: 512 + 256 ch -> 129x128x128x128)", (h.numel() + xs.numel() + 129 * 128 * 128 * 128) * 2, ms))
    if prof:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    out = [{"kernel": n, "algorithmic_bytes": b, "ms": ms, "gbs": b / ms / 1e6, "frac_of_measured_hbm": b / ms / 1e6 / peak, "peak_gbs": peak}
           for n, b, ms in rows]
    print(json.dumps(out))
    for r in out:
        print(f"{r['kernel']:75s} {r['ms']*1e3:9.1f} us  {r['gbs']:7.0f} GB/s  {100*r['frac_of_measured_hbm']:5.1f} % of {peak:.0f}", file=sys.stderr)


if __name__ == "__main__":
    main()
