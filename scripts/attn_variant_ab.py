#!/usr/bin/env python
"""A/B of the attention kernel forms (B200_ATTN_2CTA is read once per process: 1 = CTA pairs, 0 = 1-CTA kernel): correctness on
small / ragged / cross shapes vs fp32 math, then (argument `bench`) the Wan self- and cross-attention launches (40 heads x 75600^2,
40 x 75600 x 512) timed with CUDA events next to torch SDPA (cuDNN) on the same box."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from apex_studio_b200 import ops  # noqa: E402


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    res = {"variant": os.environ.get("B200_ATTN_VARIANT", "default"), "pair": os.environ.get("B200_ATTN_2CTA", "default")}
    for (B, H, Sq, Sk) in [(1, 2, 300, 333), (2, 3, 1000, 777), (1, 32, 1024, 1024), (1, 4, 2048, 512), (1, 2, 128, 64), (1, 1, 4000, 4100),
                            (1, 40, 2300, 333), (2, 24, 1500, 130)]:   # > 148 work items: the persistent grid
        torch.manual_seed(42)
        q, k, v = (torch.randn(B, H, s, 128, device="cuda", dtype=torch.bfloat16) for s in (Sq, Sk, Sk))
        out = ops.attention(q, k, v)
        ref = torch.softmax((q.float() @ k.float().transpose(-1, -2)) / math.sqrt(128), dim=-1) @ v.float()
        res[f"rel_{B}x{H}x{Sq}x{Sk}"] = round(rel(out, ref), 5)
        res["nan"] = res.get("nan", False) or bool(torch.isnan(out.float()).any())
    if len(sys.argv) > 1 and sys.argv[1] == "bench":
        sdpa = torch.nn.functional.scaled_dot_product_attention
        q, k, v = (torch.randn(1, 40, 75600, 128, device="cuda", dtype=torch.bfloat16) for _ in range(3))
        out = ops.attention(q, k, v)
        torch.cuda.synchronize()
        ref = sdpa(q[:, :2], k[:, :2], v[:, :2])
        res["rel_full_2heads_vs_sdpa"] = round(rel(out[:, :2], ref), 5)
        # flux / qwen: the joint self-attention of FLUX.1-dev 1024^2 (24 heads x 4608) and QwenImage-Edit (24 x 8704)
        for name, hs, sq, sk, iters in (("self40", 40, 75600, 75600, 8), ("self2", 2, 75600, 75600, 20),
                                        ("cross40", 40, 75600, 512, 50), ("flux24", 24, 4608, 4608, 50),
                                        ("qwen24", 24, 8704, 8704, 50)):
            qq, kk, vv = q[:, :hs, :sq], k[:, :hs, :sk], v[:, :hs, :sk]
            oo = torch.empty_like(qq)
            flop = 4.0 * hs * sq * sk * 128
            ms = timed(lambda: ops.attention(qq, kk, vv, out=oo), iters)
            ms_ref = timed(lambda: sdpa(qq, kk, vv), iters)
            res[name] = {"ms": round(ms, 3), "tflops": round(flop / ms / 1e9, 1), "sdpa_ms": round(ms_ref, 3),
                         "sdpa_tflops": round(flop / ms_ref / 1e9, 1)}
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
