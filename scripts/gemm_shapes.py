#!/usr/bin/env python
"""Per-shape timing of b200_linear vs cuBLAS (torch.matmul) for the GEMM shapes of the image models (isolated, back-to-back
launches of one shape, CUDA events).  Evidence for where the 0.75-of-peak of the Flux / QwenImage steps goes."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from apex_studio_b200 import ops  # noqa: E402

SHAPES = [  # (name, M, N, K, epilogue)
    ("flux img qkv", 4096, 9216, 3072, 0), ("flux img out", 4096, 3072, 3072, 2), ("flux img ff1", 4096, 12288, 3072, 1),
    ("flux img ff2", 4096, 3072, 12288, 2), ("flux txt qkv", 512, 9216, 3072, 0), ("flux txt out", 512, 3072, 3072, 2),
    ("flux txt ff1", 512, 12288, 3072, 1), ("flux txt ff2", 512, 3072, 12288, 2), ("flux single qkv", 4608, 9216, 3072, 0),
    ("flux single mlp", 4608, 12288, 3072, 1), ("flux single out", 4608, 3072, 15360, 2), ("qwen img qkv", 8192, 9216, 3072, 0),
    ("qwen img ff2", 8192, 3072, 12288, 2), ("wan qkv", 75600, 15360, 5120, 0), ("wan out", 75600, 5120, 5120, 2),
]


def timeit(fn, iters):
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
    dev = "cuda"
    rows = []
    for name, M, N, K, epi in SHAPES:
        x = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
        b = torch.randn(N, device=dev).bfloat16()
        g = torch.randn(N, device=dev).bfloat16()
        out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        iters = 20 if M > 20000 else 200
        ours = timeit(lambda: ops.linear(x, w, b, epilogue=epi, out=out, gate=g if epi == 2 else None), iters)
        ref = timeit(lambda: torch.matmul(x, w.t(), out=out), iters)
        fl = 2.0 * M * N * K
        rows.append({"shape": name, "M": M, "N": N, "K": K, "epi": epi, "ours_us": ours * 1e3, "cublas_us": ref * 1e3,
                     "ours_tflops": fl / ours / 1e9, "cublas_tflops": fl / ref / 1e9, "tiles": ((M + 127) // 128) * ((N + 255) // 256)})
        del x, w, out
    print(json.dumps(rows))
    for r in rows:
        print(f"{r['shape']:18s} {r['M']:6d}x{r['N']:6d}x{r['K']:6d} tiles {r['tiles']:5d}  ours {r['ours_us']:8.1f} us {r['ours_tflops']:7.1f} TF/s"
              f"   cuBLAS (no epilogue) {r['cublas_us']:8.1f} us {r['cublas_tflops']:7.1f} TF/s", file=sys.stderr)


if __name__ == "__main__":
    main()
