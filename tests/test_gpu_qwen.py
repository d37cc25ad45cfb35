"""GPU parity tests of the QwenImage path (BASELINE.json configs[2]) through the C ABI, against the CPU
oracle (oracle/qwen_dit.py, pinned bit-exactly to the reference's own model) and the reference's golden vectors.  Bars as in
tests/test_gpu_flux.py."""
import os

import pytest
import torch
import torch.nn.functional as F

import qwen_dit
from test_gpu_parity import _ulp_report, rel_l2
from test_oracle_qwen import CONFIGS, _product, inputs, kw, load

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from apex_studio_b200 import ops as _ops

    return _ops


@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("rows,dim", [(9, 48), (300, 3584)])
def test_rmsnorm_rows_vs_reference_rounding(ops, mode, rows, dim):
    torch.manual_seed(rows + mode)
    x = (torch.randn(rows, dim) * 3).bfloat16()
    w = (1 + 0.1 * torch.randn(dim)).bfloat16()
    if mode == 1:
        ref = F.rms_norm(x, (dim,), w, 1e-6)
    elif mode == 2:
        import wan_dit
        ref = wan_dit.rms_norm_across_heads(x, w, 1e-6)
    else:
        ref = qwen_dit.rms_norm(x, w, 1e-6)
    out = ops.rmsnorm_rows(x.to(DEV), w.to(DEV), 1e-6, mode).cpu()
    frac, ulps = _ulp_report(out, ref)
    assert frac <= 2e-3 and ulps <= 1.01, (mode, frac, ulps)


@pytest.mark.parametrize("epi,fn", [("EPI_SILU", F.silu), ("EPI_GELU_ERF", F.gelu)])
def test_linear_silu_and_exact_gelu_epilogues(ops, epi, fn):
    """FeedForward("linear-silu") / nn.GELU() of the HunyuanVideo-1.5 condition embedders as GEMM epilogues: rel-L2 <= 4e-3 vs
    fp32 math of the same bf16 operands."""
    torch.manual_seed(11)
    x, w, b = torch.randn(300, 264).bfloat16(), (torch.randn(520, 264) * 0.1).bfloat16(), torch.randn(520).bfloat16()
    ref = fn(x.float() @ w.float().t() + b.float())
    out = ops.linear(x.to(DEV), w.to(DEV), b.to(DEV), epilogue=getattr(ops, epi))
    assert rel_l2(out, ref) <= 4e-3


def _model(cfg, w32):
    m = _product(cfg)
    m.load_state_dict(w32, device=DEV)
    return m


@pytest.mark.parametrize("name", list(CONFIGS))
def test_qwen_forward_vs_reference_golden(name):
    cfg, g = CONFIGS[name], load(name)
    w32 = qwen_dit.make_weights(**cfg, seed=1234, dtype=torch.float32)
    m = _model(cfg, w32)
    x, enc, t, shapes, n_txt = inputs(g, torch.bfloat16)
    call = lambda: m(hidden_states=x.to(DEV), encoder_hidden_states=enc.to(DEV), encoder_hidden_states_mask=torch.ones(1, n_txt),
                     timestep=t.to(DEV), img_shapes=[shapes], txt_seq_lens=[n_txt], return_dict=False)[0]
    out = call()
    exact, ref16 = torch.from_numpy(g["out_fp32"]), torch.from_numpy(g["out_bf16"])
    assert out.dtype == torch.bfloat16 and tuple(out.shape) == tuple(exact.shape)
    ours, theirs = rel_l2(out, exact), rel_l2(ref16, exact)
    assert ours <= max(1e-3, 1.5 * theirs), (ours, theirs)
    assert rel_l2(out, ref16) <= 2e-2
    assert torch.equal(out, call())


def test_qwen_edit_width_one_block_vs_exact_oracle():
    """Qwen-Image-Edit-2509 widths (d = 3072, 24 x 128 heads, text 3584) with ONE block on a 32x32 noisy-latent grid plus one
    32x32 reference image (2048 image tokens) and 200 text tokens."""
    cfg = dict(dim=3072, heads=24, num_layers=1, in_channels=64, out_channels=16, joint_dim=3584)
    w32 = qwen_dit.make_weights(**cfg, seed=5, dtype=torch.float32, std=0.02)
    from apex_studio_b200.qwenimage import QwenImageConfig, QwenImageTransformer2DModel

    m = QwenImageTransformer2DModel(QwenImageConfig(num_layers=1))
    m.load_state_dict(w32, device=DEV)
    gen = torch.Generator().manual_seed(42)
    shapes, n_txt = [(1, 32, 32), (1, 32, 32)], 200
    x, enc, t = torch.randn(1, 2048, 64, generator=gen), torch.randn(1, n_txt, 3584, generator=gen), torch.tensor([0.5])
    out = m(hidden_states=x.to(DEV), encoder_hidden_states=enc.to(DEV), timestep=t.to(DEV), img_shapes=[shapes],
            txt_seq_lens=[n_txt], return_dict=False)[0]
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    k = dict(heads=24, num_layers=1)
    exact = qwen_dit.qwen_forward(x, enc, t, shapes, n_txt, w32, **k)
    bf = qwen_dit.qwen_forward(x.bfloat16(), enc.bfloat16(), t.bfloat16(), shapes, n_txt, {kk: v.bfloat16() for kk, v in w32.items()}, **k)
    ours, theirs = rel_l2(out, exact), rel_l2(bf, exact)
    assert torch.isfinite(out).all() and tuple(out.shape) == (1, 2048, 64)
    assert ours <= max(1e-3, 1.5 * theirs), (ours, theirs)
