"""Pins oracle/wan_vae.py (whole-sequence formulation) to the reference's streaming AutoencoderKLWan.decode."""
import os

import numpy as np
import pytest
import torch

import wan_vae
from conftest import GOLDEN

CASES = {"tiled": (True, 3), "untiled": (False, 2)}


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "wan_vae.npz"))


@pytest.mark.parametrize("name", list(CASES))
def test_decode_fp32_vs_reference(gold, name):
    tiling, sub = CASES[name]
    w = wan_vae.make_weights(base_dim=32, seed=7)
    lat = torch.from_numpy(gold[name + "_latents"])
    with torch.no_grad():
        y = wan_vae.decode(wan_vae.denormalize_latents(lat), w, use_tiling=tiling)
    assert list(y.shape) == gold[name + "_shape"].tolist()
    ref = torch.from_numpy(gold[name + "_out_fp32"])
    # streaming (reference) vs whole-sequence (oracle) convs differ only in fp32 accumulation order
    assert (y[..., ::sub, ::sub] - ref).abs().max().item() <= 5e-5
    assert y.abs().max().item() <= 1.0


def test_output_frame_count_and_causality():
    """T latent frames -> 1 + 4 (T-1) frames; frame f of the output depends only on latent frames <= ceil(f/4)."""
    w = wan_vae.make_weights(base_dim=32, seed=7)
    z = torch.randn(1, 16, 3, 6, 6, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        full = wan_vae.decoder_forward(z, w)
        assert full.shape == (1, 3, 9, 48, 48)
        first = wan_vae.decoder_forward(z[:, :, :1], w)
        two = wan_vae.decoder_forward(z[:, :, :2], w)
    assert first.shape[2] == 1 and two.shape[2] == 5
    assert torch.allclose(full[:, :, :1], first, atol=1e-5) and torch.allclose(full[:, :, :5], two, atol=1e-5)


def test_tile_grid_matches_reference_loop():
    tiles = wan_vae.tile_grid(90, 160)
    assert len(tiles) == 28 and tiles[0] == (0, 0, 32, 32) and tiles[-1] == (72, 144, 18, 16)
    assert [t[0] for t in tiles[::7]] == [0, 24, 48, 72]


def test_blend_is_in_place_and_ordered():
    a = torch.ones(1, 1, 1, 8, 8)
    b = torch.zeros(1, 1, 1, 8, 8)
    r = wan_vae._blend_v(a, b, 4)
    assert r is b and torch.allclose(b[0, 0, 0, :4, 0], torch.tensor([1.0, 0.75, 0.5, 0.25]))
