#!/usr/bin/env python
"""The reference's OWN `WanTransformerBlock` (unmodified files under baseline/_ref, see baseline/make_ref.py) run on the B200 at
the BASELINE shape -- d = 5120, 40 heads, ffn 13,824, 512 text tokens, S = 75,600 tokens, bf16, `torch.inference_mode()` like
the engine (registry.py:196) -- with the reference's stock attention backends.  This is the "reference on 1 x B200" column of
BASELINE.md section 4: torch SDPA (cuDNN fused attention) + cuBLAS + ATen elementwise kernels.

    python baseline/ref_gpu_block.py [--tokens 75600] [--iters 3] [--backends sdpa,flash]

Prints one JSON line: ms per block-forward per backend and the denoise step extrapolated from it (x 40 layers x 2 forwards;
embedders / head / scheduler excluded, < 0.1 % of the FLOPs).  Nothing of this repository's kernels is on this path."""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=75600)
    ap.add_argument("--grid", default="21,45,80")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--backends", default="sdpa,flash")
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(REF, "apps", "api", "src")):
        print(json.dumps({"unavailable": "baseline/_ref not built (python baseline/make_ref.py in the container that has /root/reference)"}))
        return 0
    os.environ["APEX_REFERENCE_API"] = os.path.join(REF, "apps", "api")
    sys.path.insert(0, REF)
    import torch
    from ref_import import bootstrap

    model_mod = bootstrap.ref("src.transformer.wan.base.model")
    attn_mod = bootstrap.ref("src.attention.functions")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    dim, heads, ffn, L = 5120, 40, 13824, 512
    f, h, w = (int(x) for x in args.grid.split(","))
    S = f * h * w
    assert S == args.tokens, (S, args.tokens)
    torch.manual_seed(1234)
    with torch.device(dev):
        block = model_mod.WanTransformerBlock(dim, ffn, heads, "rms_norm_across_heads", True, 1e-6).to(torch.bfloat16).eval()
    for p in block.parameters():
        if p.dim() >= 2 and p.shape[-1] >= 64 and p.shape[0] != 1:
            p.data.normal_(0, 0.02)
    rope = model_mod.WanRotaryPosEmbed(128, (1, 2, 2), 1024)
    lat = torch.empty(1, 16, f, 2 * h, 2 * w, device=dev)
    freqs = rope(lat)                                                   # complex [1, 1, S, 64]
    hid0 = torch.randn(1, S, dim, device=dev).bfloat16()
    ctx = torch.randn(1, L, dim, device=dev).bfloat16()
    temb = (torch.randn(1, 6, dim, device=dev) * 0.5).bfloat16()
    flops_block = 8.0 * S * dim * dim + 4.0 * S * S * dim + (4.0 * S * dim * dim + 4.0 * L * dim * dim + 4.0 * S * L * dim) + 4.0 * S * dim * ffn
    res = {"what": "reference WanTransformerBlock on this GPU (unmodified reference files, torch " + torch.__version__ + ")",
           "tokens": S, "gpu": torch.cuda.get_device_name(0), "backends": {}}
    for name in args.backends.split(","):
        try:
            if not attn_mod.attention_register.is_available(name):
                res["backends"][name] = {"unavailable": "not registered / not available in this image"}
                continue
            attn_mod.attention_register.set_default(name)
            with torch.inference_mode():
                for _ in range(2):
                    hs = hid0.clone()
                    out = block(hs, ctx, temb, freqs)
                torch.cuda.synchronize()
                times = []
                for _ in range(args.iters):
                    hs = hid0.clone()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    out = block(hs, ctx, temb, freqs)
                    e1.record()
                    torch.cuda.synchronize()
                    times.append(e0.elapsed_time(e1))
            ms = sum(times) / len(times)
            res["backends"][name] = {"ms_per_block": ms, "tflops": flops_block / (ms * 1e-3) / 1e12,
                                     "steps_per_sec_extrapolated": 1.0 / (ms * 1e-3 * 40 * 2), "finite": bool(torch.isfinite(out.float()).all()),
                                     "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
        except Exception as e:  # noqa
            res["backends"][name] = {"error": repr(e)[:300]}
    ok = {k: v for k, v in res["backends"].items() if "ms_per_block" in v}
    if ok:
        best = min(ok, key=lambda k: ok[k]["ms_per_block"])
        res.update(backend=best, ms_per_block=ok[best]["ms_per_block"], steps_per_sec_extrapolated=ok[best]["steps_per_sec_extrapolated"],
                   extrapolation="x 40 layers x 2 forwards; embedders / head / scheduler excluded")
    print(json.dumps(res), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
