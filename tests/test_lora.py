"""LoRA hooks: host-side key handling on CPU, merge numerics on the GPU.

Reference: apps/api/src/lora/manager.py:454-606 (load_into -> load_lora_adapter + set_adapters), :383-396, :812-838;
lora_converter.py:139-163.  Numerics oracle: oracle/lora.py (PEFT runtime form, parity unpinned -- see its header)."""
import numpy as np
import pytest
import torch

import lora as lora_oracle
import wan_dit
from apex_studio_b200 import lora as L

DEV = "cuda"
CFG = dict(dim=256, heads=2, ffn_dim=512, num_layers=2, text_dim=64, freq_dim=256)


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


# ------------------------------------------------------------------------------------------------ host logic (CPU)
def test_strip_adapter_name_and_prefix_detection():
    st = {"transformer.blocks.0.attn1.to_q.lora_A.default.weight": torch.zeros(4, 8),
          "transformer.blocks.0.attn1.to_q.lora_B.default.weight": torch.zeros(8, 4)}
    out = L.strip_adapter_name_from_keys(st)
    assert sorted(out) == ["transformer.blocks.0.attn1.to_q.lora_A.weight", "transformer.blocks.0.attn1.to_q.lora_B.weight"]
    assert L.get_prefix_key(list(out)) == "transformer"
    assert L.get_prefix_key(["blocks.0.x", "transformer.y"]) is None          # first AND last key must carry it
    assert L.get_prefix_key(["diffusion_model.a", "diffusion_model.b"]) == "diffusion_model"
    norm = L.normalize_lora_state_dict(st)
    assert sorted(norm) == ["blocks.0.attn1.to_q.lora_A.weight", "blocks.0.attn1.to_q.lora_B.weight"]


def test_alpha_is_folded_like_the_converter():
    A, B = torch.ones(4, 8), torch.ones(8, 4)
    st = {"m.lora_down.weight": A, "m.lora_up.weight": B, "m.alpha": torch.tensor(1.0)}
    out = L.normalize_lora_state_dict(st, prefix=None)
    assert sorted(out) == ["m.lora_A.weight", "m.lora_B.weight"]
    # alpha / rank = 0.25 -> scale_down 0.25 doubles while 2*down < up: (0.25, 1) -> (0.5, 0.5)
    assert torch.allclose(out["m.lora_A.weight"], A * 0.5) and torch.allclose(out["m.lora_B.weight"], B * 0.5)
    prod = out["m.lora_B.weight"] @ out["m.lora_A.weight"]
    assert torch.allclose(prod, (B @ A) * 0.25)


def test_split_modules_validates_shapes():
    with pytest.raises(ValueError):
        L.split_modules({"m.lora_A.weight": torch.zeros(4, 8)})
    with pytest.raises(ValueError):
        L.split_modules({"m.lora_A.weight": torch.zeros(4, 8), "m.lora_B.weight": torch.zeros(8, 3)})
    with pytest.raises(ValueError):
        L.normalize_lora_state_dict({"m.lora_magnitude_vector": torch.zeros(3)}, prefix=None)


def test_lora_target_resolves_fused_projections():
    from apex_studio_b200.wan import WanConfig, WanTransformer3DModel

    m = WanTransformer3DModel(WanConfig(num_attention_heads=2, text_dim=64, ffn_dim=512, num_layers=1))
    m.load_state_dict(wan_dit.make_weights(**dict(CFG, num_layers=1), seed=1, dtype=torch.float32), device="cpu")
    assert m.lora_target("blocks.0.attn1.to_k") == ("blocks.0.attn1.to_qkv.weight", 256, 256, "blocks.0.attn1.to_qkv.bias")
    assert m.lora_target("blocks.0.attn2.to_v") == ("blocks.0.attn2.to_kv.weight", 256, 256, "blocks.0.attn2.to_kv.bias")
    assert m.lora_target("blocks.0.attn2.to_q")[:3] == ("blocks.0.attn2.to_q.weight", 0, 256)
    assert m.lora_target("blocks.0.ffn.net.0.proj")[:3] == ("blocks.0.ffn.net.0.proj.weight", 0, 512)
    with pytest.raises(ValueError):
        m.lora_target("blocks.0.attn1.nope")
    with pytest.raises(ValueError):
        m.lora_target("patch_embedding")
    # the PeftAdapterMixin surface the manager probes (manager.py:470-486)
    assert hasattr(m, "set_adapters") and hasattr(m, "load_lora_adapter")
    w = wan_dit.make_weights(**dict(CFG, num_layers=1), seed=1, dtype=torch.float32)
    lo = lora_oracle.make_lora(w, ["blocks.0.attn1.to_q"], rank=4)
    m.load_lora_adapter({"transformer." + k: v for k, v in lo.items()}, adapter_name="a", prefix="transformer")
    assert list(m.peft_config) == ["a"] and m.peft_config["a"]["r"] == 4
    with pytest.raises(ValueError):
        m.load_lora_adapter(lo, adapter_name="a", prefix=None)            # duplicate adapter name
    with pytest.raises(ValueError):
        m.set_adapters(["missing"], [1.0])
    with pytest.raises(ValueError):
        m.set_adapters(["a"], [1.0, 2.0])
    with pytest.raises(ValueError):                                        # no CPU fallback: the merge is a CUDA GEMM
        m.set_adapters(["a"], [1.0])


def test_mixed_rank_scaling_follows_the_managers_peft_config():
    """manager.py:433-447: r = lora_alpha = the modal rank, rank_pattern for the rest -> PEFT scales a module of rank r_m
    by r_modal / r_m.  Host logic only (no GPU)."""
    from apex_studio_b200.lora import peft_scaling

    ranks = {"a.to_q": 16, "a.to_k": 16, "a.to_v": 4, "b.ffn": 32, "b.to_out.0": 16}
    sc, cfg = peft_scaling(ranks)
    assert cfg["r"] == 16 and cfg["lora_alpha"] == 16 and cfg["rank_pattern"] == {"a.to_v": 4, "b.ffn": 32}
    assert sc == {"a.to_q": 1.0, "a.to_k": 1.0, "a.to_v": 4.0, "b.ffn": 0.5, "b.to_out.0": 1.0}
    assert sc == lora_oracle.manager_scaling(ranks)
    # explicit metadata (what the manager passes to load_lora_adapter) wins, incl. alpha_pattern and suffix matching
    sc2, _ = peft_scaling(ranks, {"r": 16, "lora_alpha": 8, "rank_pattern": {"to_v": 4, "b.ffn": 32}, "alpha_pattern": {"b.ffn": 64}})
    assert sc2 == {"a.to_q": 0.5, "a.to_k": 0.5, "a.to_v": 2.0, "b.ffn": 2.0, "b.to_out.0": 0.5}
    with pytest.raises(ValueError):
        peft_scaling(ranks, {"r": 16, "lora_alpha": 16, "rank_pattern": {}})      # to_v is rank 4, metadata says 16


def test_oracle_runtime_form_equals_merged_weight():
    g = torch.Generator().manual_seed(3)
    x, W, b = torch.randn(5, 16, generator=g), torch.randn(12, 16, generator=g), torch.randn(12, generator=g)
    ads = [(torch.randn(4, 16, generator=g), torch.randn(12, 4, generator=g), 0.7),
           (torch.randn(2, 16, generator=g), torch.randn(12, 2, generator=g), -1.3)]
    y1 = lora_oracle.lora_linear_runtime(x, W, b, ads)
    y2 = torch.nn.functional.linear(x, lora_oracle.merged_weight(W, ads), b)
    assert torch.allclose(y1, y2, atol=1e-5)


# ------------------------------------------------------------------------------------------------ GPU
def _model():
    from apex_studio_b200.wan import WanConfig, WanTransformer3DModel

    w32 = wan_dit.make_weights(**CFG, seed=1234, dtype=torch.float32)
    m = WanTransformer3DModel(WanConfig(num_attention_heads=2, text_dim=64, ffn_dim=512, num_layers=2))
    m.load_state_dict(w32, device=DEV)
    return m, w32


MODULES = ["blocks.0.attn1.to_q", "blocks.0.attn1.to_k", "blocks.0.attn1.to_v", "blocks.0.attn1.to_out.0",
           "blocks.1.attn2.to_q", "blocks.1.attn2.to_k", "blocks.1.attn2.to_v", "blocks.1.ffn.net.0.proj",
           "blocks.1.ffn.net.2", "condition_embedder.time_proj"]


@pytest.mark.gpu
@pytest.mark.parametrize("rank", [4, 16, 36, 128])
def test_merge_equals_fp32_math_and_unmerge_is_bit_exact(rank):
    m, w32 = _model()
    base = {k: v.clone() for k, v in m.w.items()}
    lo = lora_oracle.make_lora(w32, MODULES, rank=rank, seed=rank)
    m.load_lora_adapter(lo, adapter_name="x", prefix=None)
    m.set_adapters(["x"], [0.8])
    for mod in MODULES:
        wkey, r0, rows, _ = m.lora_target(mod)
        got = m.w[wkey][r0:r0 + rows].float().cpu()
        # same operands the kernel saw: bf16 base, bf16(0.8 * B), bf16 A; fp32 accumulate; one rounding
        A16, B16 = lo[mod + ".lora_A.weight"].bfloat16().float(), (lo[mod + ".lora_B.weight"] * 0.8).bfloat16().float()
        want = (base[wkey][r0:r0 + rows].float().cpu() + B16 @ A16)
        ulp = (want.abs().clamp_min(2.0 ** -126)).log2().floor().exp2() * 2.0 ** -7
        # correctly rounded, up to the fp32 summation order: where base and delta cancel (|want| << |base|) the
        # accumulation error (<= (r + 1) * 2^-24 * sum of magnitudes) is not small against ulp(want), so it is allowed for
        # explicitly (an element with want ~ 3e-8 missed the plain 1-ulp bar by 4 ulp on the GPU, rank 36)
        acc = (rank + 1) * 2.0 ** -24 * (base[wkey][r0:r0 + rows].float().cpu().abs() + B16.abs() @ A16.abs())
        err = ((got - want).abs() - acc).clamp_min(0) / ulp
        assert err.max() <= 1.0 and (err > 0.51).float().mean() < 1e-3, (mod, err.max().item())
        # and close to the exact (unrounded factors) merge
        exact = lora_oracle.merged_weight(w32[mod + ".weight"], [(lo[mod + ".lora_A.weight"], lo[mod + ".lora_B.weight"], 0.8)])
        assert rel_l2(got, exact) <= 4e-3
    # keys that share no storage with an adapted weight (the stacked layouts alias their per-layer views, e.g.
    # attn2_kv_all.weight <-> blocks.N.attn2.to_kv.weight)
    touched = {m.w[m.lora_target(mod)[0]].untyped_storage().data_ptr() for mod in MODULES}
    untouched = [k for k in base if m.w[k].untyped_storage().data_ptr() not in touched]
    assert len(untouched) > 10 and all(torch.equal(m.w[k], base[k]) for k in untouched)
    m.set_adapters(["x"], [0.0])
    assert all(torch.equal(m.w[k], base[k]) for k in base)
    m.set_adapters(["x"], [1.0])
    m.delete_adapters("x")
    assert all(torch.equal(m.w[k], base[k]) for k in base) and m.peft_config == {}


@pytest.mark.gpu
def test_two_adapters_forward_vs_oracle_runtime_form():
    """DiT forward with two active adapters vs the exact-math oracle with the adapters merged in fp32 (== PEFT's
    runtime form by linearity).  Tolerance: the same bar as the un-adapted forward (rel-L2 <= max(1e-3, 1.5 x the bf16
    error of the oracle run in bf16))."""
    import os

    from conftest import GOLDEN

    m, w32 = _model()
    g = np.load(os.path.join(GOLDEN, "dit_s72.npz"))
    lat, t, text = torch.from_numpy(g["latents"]), torch.from_numpy(g["timestep"]), torch.from_numpy(g["text"])
    mods = [f"blocks.{i}.{n}" for i in range(2) for n in ("attn1.to_q", "attn1.to_k", "attn1.to_v", "attn1.to_out.0",
                                                           "attn2.to_q", "attn2.to_k", "attn2.to_v", "attn2.to_out.0",
                                                           "ffn.net.0.proj", "ffn.net.2")]
    lo1 = lora_oracle.make_lora(w32, mods, rank=16, seed=1)
    lo2 = lora_oracle.make_lora(w32, mods[::2], rank=8, seed=2)
    m.load_lora_adapter({"transformer." + k: v for k, v in lo1.items()}, adapter_name="lightning", prefix="transformer")
    m.load_lora_adapter({k.replace(".weight", ".default.weight"): v for k, v in lo2.items()}, adapter_name="style", prefix=None)
    m.set_adapters(["lightning", "style"], [1.0, 0.6])
    assert m.active_adapters() == ["lightning", "style"]
    out = m(lat.to(DEV, torch.bfloat16), t.to(DEV), text.to(DEV, torch.bfloat16), return_dict=False)[0]
    wm = lora_oracle.merge_into_state_dict(lora_oracle.merge_into_state_dict(w32, lo1, 1.0), lo2, 0.6)
    kw = dict(heads=2, num_layers=2, freq_dim=256)
    exact = wan_dit.dit_forward(lat, t, text, wm, **kw)
    bf = wan_dit.dit_forward(lat.bfloat16(), t, text.bfloat16(), {k: v.bfloat16() for k, v in wm.items()}, **kw)
    base_out = wan_dit.dit_forward(lat, t, text, w32, **kw)
    assert rel_l2(exact, base_out) > 3e-2                       # the adapters really change the output
    assert rel_l2(out, exact) <= max(1e-3, 1.5 * rel_l2(bf, exact)), (rel_l2(out, exact), rel_l2(bf, exact))
    m.disable_lora()
    out0 = m(lat.to(DEV, torch.bfloat16), t.to(DEV), text.to(DEV, torch.bfloat16), return_dict=False)[0]
    assert rel_l2(out0, torch.from_numpy(g["out_bf16"])) <= 2e-2


@pytest.mark.gpu
def test_mixed_rank_adapter_merge_uses_per_module_scaling():
    """One adapter file with ranks 16 / 16 / 4 (modal rank 16): the rank-4 module must be merged with 4x the weight
    (PEFT: lora_alpha / r_m = 16 / 4), the rank-16 modules with 1x."""
    m, w32 = _model()
    mods = ["blocks.0.attn1.to_q", "blocks.0.attn1.to_k", "blocks.0.attn1.to_v"]
    lo = {}
    for mod, r in zip(mods, (16, 16, 4)):
        lo.update(lora_oracle.make_lora(w32, [mod], rank=r, seed=r + len(mod)))
    m.load_lora_adapter(lo, adapter_name="mixed", prefix=None)
    assert m.peft_config["mixed"]["r"] == 16 and m.peft_config["mixed"]["rank_pattern"] == {"blocks.0.attn1.to_v": 4}
    m.set_adapters(["mixed"], [0.5])
    sc = lora_oracle.manager_scaling({mod: r for mod, r in zip(mods, (16, 16, 4))})
    for mod in mods:
        wkey, r0, rows, _ = m.lora_target(mod)
        got = m.w[wkey][r0:r0 + rows].float().cpu()
        exact = lora_oracle.merged_weight(w32[mod + ".weight"], [(lo[mod + ".lora_A.weight"], lo[mod + ".lora_B.weight"], 0.5 * sc[mod])])
        wrong = lora_oracle.merged_weight(w32[mod + ".weight"], [(lo[mod + ".lora_A.weight"], lo[mod + ".lora_B.weight"], 0.5)])
        assert rel_l2(got, exact) <= 4e-3
        if sc[mod] != 1.0:
            assert rel_l2(got, wrong) > 2e-2


@pytest.mark.gpu
def test_merged_bf16_weight_vs_peft_runtime_form_small_delta():
    """Quantifies what merging costs against PEFT's runtime form  y = W x + s B (A x)  for a SMALL-delta adapter (the
    regime where rounding W + s B A to bf16 loses most of the delta): relative L2 of the adapter's *contribution*
    (y - W x) reproduced by the merged weight, next to the same figure for the runtime form evaluated in bf16 as PEFT does
    (A x, B(.), the scale and the sum each round the activation).  Expected from the CPU emulation of both forms: |delta| ~
    2^-9 |W| (half a bf16 ulp of the weight): merged 0.61 vs runtime 1.05; |delta| ~ 2^-5 |W|: merged 0.052 vs runtime 0.074;
    output-level error merged 1.2e-3 / 1.7e-3 vs runtime 2.1e-3 / 2.3e-3.  Stated bars: the merged form is at least as close
    to exact as the runtime form on both measures, contribution error <= 0.7 (tiny) / 0.06 (small), output error <= 2.5e-3."""
    torch.manual_seed(0)
    N, K, r = 512, 512, 16
    W = (torch.randn(N, K) * 0.02).bfloat16()
    x = torch.randn(256, K).bfloat16()
    from apex_studio_b200 import ops
    report = {}
    for name, rel_delta in (("tiny", 2.0 ** -9), ("small", 2.0 ** -5)):
        A, B = torch.randn(r, K) * 0.05, torch.randn(N, r) * 0.05
        delta = B @ A
        s = rel_delta * W.float().norm() / delta.norm()
        exact_contrib = x.float() @ (s * delta).t()
        w_eff = W.to(DEV).clone()
        ops.linear((B * s).bfloat16().to(DEV).contiguous(), A.bfloat16().t().contiguous().to(DEV), None, epilogue=ops.EPI_GATE_RES,
                   out=w_eff, gate=None)
        y_merged = x.float() @ w_eff.float().cpu().t()
        y_base = x.float() @ W.float().t()
        # PEFT runtime form in bf16 (each op rounds)
        xa = torch.nn.functional.linear(x, A.bfloat16())
        y_rt = (torch.nn.functional.linear(x, W) + torch.nn.functional.linear(xa, B.bfloat16()) * s.bfloat16()).float()
        exact = y_base + exact_contrib
        report[name] = dict(contrib_err_merged=rel_l2(y_merged - y_base, exact_contrib), contrib_err_runtime=rel_l2(y_rt - y_base, exact_contrib),
                            out_err_merged=rel_l2(y_merged, exact), out_err_runtime=rel_l2(y_rt, exact))
    print("[lora merged vs runtime]", report)
    assert report["tiny"]["contrib_err_merged"] <= 0.7 and report["small"]["contrib_err_merged"] <= 0.06, report
    for v in report.values():
        assert v["contrib_err_merged"] <= v["contrib_err_runtime"] and v["out_err_merged"] <= v["out_err_runtime"], report
        assert v["out_err_merged"] <= 2.5e-3, report
