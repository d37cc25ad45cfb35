"""Pins the HunyuanVideo-1.5 VAE-decode CPU oracle (oracle/hy15_vae.py) to the reference: tests/golden/hy15_vae.npz was
produced by the reference's OWN AutoencoderKLHunyuanVideo15 (oracle/make_golden.py golden_hy15vae)."""
import os

import numpy as np
import pytest
import torch

import hy15_vae
from conftest import GOLDEN

CH = (128, 128, 64, 64, 32)
CASES = {"untiled": (False, 1), "tiled": (True, 2)}


def load():
    return np.load(os.path.join(GOLDEN, "hy15_vae.npz"))


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("tag,dt", [("fp32", torch.float32), ("bf16", torch.bfloat16)])
def test_decode_vs_reference(name, tag, dt):
    g = load()
    tiling, sub = CASES[name]
    w = hy15_vae.make_weights(CH, seed=7, dtype=dt)
    z = torch.from_numpy(g[name + "_latents"]).to(dt)
    y = hy15_vae.tiled_decode(z, w, CH) if tiling else hy15_vae.decoder(z, w, CH)
    assert tuple(y.shape) == tuple(g[name + "_shape"])
    ref = torch.from_numpy(g[f"{name}_out_{tag}"])
    got = y[..., ::sub, ::sub].float()
    if dt == torch.float32:
        # same ops, same order; conv3d on sliced / whole tensors may pick different CPU kernels -> allow fp32 noise
        assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5), (got - ref).abs().max().item()
    else:
        mism = (got != ref).float().mean().item()
        assert mism <= 2e-3 and (got - ref).abs().max().item() <= 0.04, (mism, (got - ref).abs().max().item())


def test_causal_mask_matches_reference_definition():
    m = hy15_vae.causal_mask(3, 4, torch.float32)
    ref = torch.full((12, 12), float("-inf"))
    for i in range(12):
        ref[i, : (i // 4 + 1) * 4] = 0
    assert torch.equal(m, ref)
