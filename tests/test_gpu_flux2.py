"""GPU parity tests of the Flux2 path (BASELINE.json configs[0] family; SURVEY.md section 8 f1 "SwiGLU MLP for Flux2") through the C
ABI, against the CPU oracle (oracle/flux2_dit.py, pinned bit-exactly to the reference's own model) and its golden vectors."""
import os

import pytest
import torch

import flux2_dit
from test_gpu_parity import rel_l2
from test_oracle_flux2 import CONFIGS, _product, inputs, kw, load

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("name", list(CONFIGS))
def test_flux2_forward_vs_reference_golden(name):
    cfg, g = CONFIGS[name], load(name)
    m = _product(cfg)
    m.load_state_dict(flux2_dit.make_weights(**cfg, seed=1234, dtype=torch.float32), device=DEV)
    x, enc, t, img_ids, txt_ids, guidance = inputs(g, torch.bfloat16)
    call = lambda: m(x.to(DEV), enc.to(DEV), t.to(DEV), img_ids, txt_ids, None if guidance is None else guidance.to(DEV),
                     return_dict=False)[0]
    out = call()
    exact, ref16 = torch.from_numpy(g["out_fp32"]), torch.from_numpy(g["out_bf16"])
    assert out.dtype == torch.bfloat16 and tuple(out.shape) == tuple(exact.shape)
    ours, theirs = rel_l2(out, exact), rel_l2(ref16, exact)
    assert ours <= max(1e-3, 1.5 * theirs), (ours, theirs)
    assert rel_l2(out, ref16) <= 2e-2
    assert torch.equal(out, call())


def test_flux2_klein_width_blocks_vs_exact_oracle():
    """Klein-4B-like widths (d = 3072 = 24 x 128 heads, mlp 9216, 128 latent channels, text width 7680) with one dual and one
    parallel single block on a 256x256 image (16 x 16 = 256 latent tokens) + 512 text tokens -- BASELINE configs[0]'s shape."""
    cfg = dict(dim=3072, heads=24, num_layers=1, num_single_layers=1, in_channels=128, joint_dim=7680, guidance_embeds=False)
    w32 = flux2_dit.make_weights(**cfg, seed=5, dtype=torch.float32, std=0.02)
    m = _product(cfg)
    m.load_state_dict(w32, device=DEV)
    gen = torch.Generator().manual_seed(42)
    x, enc, t = torch.randn(1, 256, 128, generator=gen), torch.randn(1, 512, 7680, generator=gen), torch.tensor([0.5])
    img_ids = torch.zeros(256, 4)
    img_ids[:, 1], img_ids[:, 2] = torch.arange(256) // 16, torch.arange(256) % 16
    txt_ids = torch.zeros(512, 4)
    txt_ids[:, 3] = torch.arange(512)
    out = m(x.to(DEV), enc.to(DEV), t.to(DEV), img_ids, txt_ids, None, return_dict=False)[0]
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    k = kw(cfg)
    exact = flux2_dit.flux2_forward(x, enc, t, img_ids, txt_ids, None, w32, **k)
    bf = flux2_dit.flux2_forward(x.bfloat16(), enc.bfloat16(), t.bfloat16(), img_ids, txt_ids, None,
                                 {kk: v.bfloat16() for kk, v in w32.items()}, **k)
    ours, theirs = rel_l2(out, exact), rel_l2(bf, exact)
    assert torch.isfinite(out).all() and tuple(out.shape) == (1, 256, 128)
    assert ours <= max(1e-3, 1.5 * theirs), (ours, theirs)
