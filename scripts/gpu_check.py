"""Bring-up checks run on the GPU box (scratch tooling, not part of the test-suite).

    python scripts/gpu_check.py <test> [args]

Each test prints one JSON line.  Run every test in its own process under `timeout` so that a hung kernel
cannot take the whole call down (see scripts/gpu_run_all.sh).
"""
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from apex_studio_b200 import ops


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def timeit(fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def err_report(out, ref, name, extra=None):
    o, r = out.float(), ref.float()
    d = (o - r).abs()
    rep = {
        "test": name,
        "rel_l2": rel_l2(o, r),
        "max_abs": d.max().item(),
        "ref_max": r.abs().max().item(),
        "nan": bool(torch.isnan(o).any().item()),
    }
    if extra:
        rep.update(extra)
    return rep


def structure(out, ref, tile=(8, 8)):
    """Coarse map of where a 2-D result is wrong: fraction of bad elements per block of rows/cols."""
    d = (out.float() - ref.float()).abs() > (0.05 * ref.float().abs().max() + 1e-3)
    R, C = d.shape
    rows_bad = d.any(dim=1)
    cols_bad = d.any(dim=0)
    return {
        "bad_frac": d.float().mean().item(),
        "bad_rows_first": rows_bad.nonzero().flatten()[:16].tolist(),
        "bad_cols_first": cols_bad.nonzero().flatten()[:16].tolist(),
        "n_bad_rows": int(rows_bad.sum()),
        "n_bad_cols": int(cols_bad.sum()),
    }


# --------------------------------------------------------------------------------------------------
def test_elementwise():
    torch.manual_seed(0)
    dev = "cuda"
    rows, dim, heads = 1000, 5120, 40
    x = torch.randn(rows, dim, device=dev).bfloat16()
    scale = (torch.randn(dim, device=dev) * 0.1).bfloat16()
    shift = (torch.randn(dim, device=dev) * 0.1).bfloat16()
    # reference semantics (model.py:56-116, ops.py:37-56)
    ref = torch.nn.functional.layer_norm(x.float(), (dim,), None, None, 1e-6).to(torch.bfloat16)
    ref.addcmul_(ref, scale)
    ref.add_(shift)
    out = ops.layernorm_modulate(x, scale, shift, eps=1e-6)
    r1 = err_report(out, ref, "layernorm_modulate", {"mismatch_frac": (out != ref).float().mean().item()})
    # affine LN
    w = (1 + 0.1 * torch.randn(dim, device=dev)).bfloat16()
    b = (0.1 * torch.randn(dim, device=dev)).bfloat16()
    ref2 = torch.nn.functional.layer_norm(x.float(), (dim,), w.float(), b.float(), 1e-6).to(torch.bfloat16)
    out2 = ops.layernorm_modulate(x, ln_weight=w, ln_bias=b, eps=1e-6)
    r2 = err_report(out2, ref2, "layernorm_affine", {"mismatch_frac": (out2 != ref2).float().mean().item()})
    # rmsnorm + rope
    hd = dim // heads
    ang = torch.rand(rows, hd // 2, device=dev, dtype=torch.float64) * 6.28
    cos, sin = ang.cos().to(torch.bfloat16), ang.sin().to(torch.bfloat16)
    rope = torch.stack([cos, sin], dim=-1).reshape(rows, hd).contiguous()
    wq = (1 + 0.1 * torch.randn(dim, device=dev)).bfloat16()
    xr = x.clone()
    y = xr.float().pow(2).mean(-1, keepdim=True).add(1e-6).rsqrt()
    xr.mul_(y.to(torch.bfloat16))
    xr.mul_(wq)
    xv = xr.view(rows, heads, hd // 2, 2)
    re, im = xv[..., 0], xv[..., 1]
    c, s = cos[:, None, :], sin[:, None, :]
    re0 = re.clone()
    re.mul_(c).addcmul_(im, s.expand_as(im), value=-1.0)
    im.mul_(c).addcmul_(re0, s.expand_as(im), value=1.0)
    xo = x.clone()
    ops.rmsnorm_rope_(xo, wq, rope, heads, 1e-6)
    r3 = err_report(xo, xr, "rmsnorm_rope", {"mismatch_frac": (xo != xr).float().mean().item()})
    # gate residual
    h = torch.randn(rows, dim, device=dev).bfloat16()
    yv = torch.randn(rows, dim, device=dev).bfloat16()
    g = torch.randn(dim, device=dev).bfloat16()
    href = h.clone()
    yy = yv.clone()
    yy.mul_(g)
    href.add_(yy)
    ho = h.clone()
    ops.gate_residual_(ho, yv, g)
    r4 = err_report(ho, href, "gate_residual", {"mismatch_frac": (ho != href).float().mean().item()})
    # cfg
    cnd = torch.randn(4, 16, 5, 30, 40, device=dev).bfloat16()
    unc = torch.randn(4, 16, 5, 30, 40, device=dev).bfloat16()
    cref = (unc + 4.0 * (cnd - unc)).float()
    co = ops.cfg_combine(cnd, unc, 4.0)
    r5 = err_report(co, cref, "cfg_combine", {"mismatch_frac": (co != cref).float().mean().item()})
    # bandwidth of the big rows
    rows_b = 75600
    xb = torch.randn(rows_b, dim, device=dev).bfloat16()
    ob = torch.empty_like(xb)
    ms = timeit(lambda: ops.layernorm_modulate(xb, scale, shift, out=ob))
    r1["full_ms"] = ms
    r1["full_gbs"] = 2 * xb.numel() * 2 / ms / 1e6
    ropeb = rope[:1].expand(rows_b, hd).contiguous()
    ms = timeit(lambda: ops.rmsnorm_rope_(xb, wq, ropeb, heads, 1e-6))
    r3["full_ms"] = ms
    r3["full_gbs"] = 2 * xb.numel() * 2 / ms / 1e6
    return [r1, r2, r3, r4, r5]


def test_linear(M=256, N=512, K=256, epi=0, bench=0):
    torch.manual_seed(1)
    dev = "cuda"
    x = (torch.randn(M, K, device=dev)).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = (torch.randn(N, device=dev) * 0.5).bfloat16()
    acc = x.float() @ w.float().t() + b.float()
    extra = {"M": M, "N": N, "K": K, "epi": epi}
    if epi == 0:
        ref = acc
        out = ops.linear(x, w, b)
    elif epi == 1:
        ref = torch.nn.functional.gelu(acc, approximate="tanh")
        out = ops.linear(x, w, b, epilogue=ops.EPI_GELU_TANH)
    elif epi == 2:
        h = torch.randn(M, N, device=dev).bfloat16()
        g = torch.randn(N, device=dev).bfloat16()
        ref = h.float() + g.float() * acc
        out = h.clone()
        ops.linear(x, w, b, epilogue=ops.EPI_GATE_RES, out=out, gate=g)
    else:
        ref = acc
        out = ops.linear(x, w, b, epilogue=ops.EPI_BIAS_F32)
    torch.cuda.synchronize()
    rep = err_report(out, ref, "linear", extra)
    if rep["rel_l2"] > 2e-2 or rep["nan"]:
        rep["structure"] = structure(out, ref)
        # hypothesis probes: is the result a permutation / partial sum of the truth?
        nok = acc if epi != 1 else ref
        rep["first_row_out"] = out[0, :8].float().tolist()
        rep["first_row_ref"] = ref[0, :8].float().tolist()
    if bench:
        ms = timeit(lambda: ops.linear(x, w, b), iters=10, warmup=3)
        rep["ms"] = ms
        rep["tflops"] = 2.0 * M * N * K / ms / 1e9
    return [rep]


def test_attn(B=1, H=2, Sq=256, Sk=256, bench=0, seed=42):
    torch.manual_seed(seed)
    dev = "cuda"
    D = 128
    q = torch.randn(B, H, Sq, D, device=dev, dtype=torch.bfloat16)
    k = torch.randn(B, H, Sk, D, device=dev, dtype=torch.bfloat16)
    v = torch.randn(B, H, Sk, D, device=dev, dtype=torch.bfloat16)
    out = ops.attention(q, k, v)
    torch.cuda.synchronize()
    extra = {"B": B, "H": H, "Sq": Sq, "Sk": Sk}
    if Sq * Sk * H * B <= 64 * 4096 * 4096:
        s = (q.float() @ k.float().transpose(-1, -2)) / math.sqrt(D)
        ref = torch.softmax(s, dim=-1) @ v.float()
    else:
        ref = torch.nn.functional.scaled_dot_product_attention(q, k, v)
    rep = err_report(out, ref, "attn", extra)
    sd = torch.nn.functional.scaled_dot_product_attention(q, k, v)
    rep["sdpa_rel_l2_vs_ref"] = rel_l2(sd, ref)
    if rep["rel_l2"] > 2e-2 or rep["nan"]:
        rep["structure"] = structure(out[0, 0], ref[0, 0])
        rep["first_row_out"] = out[0, 0, 0, :8].float().tolist()
        rep["first_row_ref"] = ref[0, 0, 0, :8].float().tolist()
    if bench:
        ms = timeit(lambda: ops.attention(q, k, v), iters=3, warmup=1)
        rep["ms"] = ms
        rep["tflops"] = 4.0 * B * H * Sq * Sk * D / ms / 1e9
        ms2 = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), iters=3, warmup=1)
        rep["sdpa_ms"] = ms2
        rep["sdpa_tflops"] = 4.0 * B * H * Sq * Sk * D / ms2 / 1e9
    return [rep]


if __name__ == "__main__":
    name = sys.argv[1]
    args = [int(a) for a in sys.argv[2:]]
    t0 = time.time()
    try:
        reps = globals()["test_" + name](*args)
    except Exception as e:  # noqa
        import traceback

        reps = [{"test": name, "args": args, "exception": repr(e), "tb": traceback.format_exc()[-1500:]}]
    for r in reps:
        r["wall_s"] = round(time.time() - t0, 1)
        print(json.dumps(r), flush=True)
