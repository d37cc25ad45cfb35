"""GPU parity tests of the HunyuanVideo-1.5 path (BASELINE.json configs[4]) through the C ABI, against
the CPU oracle (oracle/hy15_dit.py, pinned bit-exactly to the reference's own model) and the reference's golden vectors.
Bar for whole forwards: relative L2 vs the exact-math fp32 oracle <= max(1e-3, 1.5 x the reference's own bf16 error
against that oracle), and <= 2e-2 vs the reference's bf16 output."""
import os

import pytest
import torch

import hy15_dit
from test_gpu_parity import rel_l2
from test_oracle_hy15 import CONFIGS, _product, inputs, kw, load

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(cfg, w32):
    m = _product(cfg)
    m.load_state_dict(w32, device=DEV)
    return m


def _call(m, x, t, text, mask, text2, mask2, img):
    return m(x.to(DEV), t.to(DEV), text.to(DEV), mask, encoder_hidden_states_2=text2.to(DEV), encoder_attention_mask_2=mask2,
             image_embeds=img.to(DEV), return_dict=False)[0]


@pytest.mark.parametrize("name", list(CONFIGS))
def test_hy15_forward_vs_reference_golden(name):
    cfg, g = CONFIGS[name], load(name)
    w32 = hy15_dit.make_weights(**cfg, seed=1234, dtype=torch.float32)
    m = _model(cfg, w32)
    args = inputs(g, torch.bfloat16)
    out = _call(m, *args)
    assert out.dtype == torch.bfloat16 and tuple(out.shape) == tuple(g["out_bf16"].shape)
    exact = hy15_dit.hy15_forward(*inputs(g, torch.float32), w32, **kw(cfg))    # exact math (no fp32 aliasing quirk)
    ref16 = torch.from_numpy(g["out_bf16"])
    ours, theirs = rel_l2(out, exact), rel_l2(ref16, exact)
    assert ours <= max(1e-3, 1.5 * theirs), (ours, theirs)
    assert rel_l2(out, ref16) <= 2e-2
    assert torch.equal(out, _call(m, *args))                                       # workspace reuse, deterministic


def test_hy15_token_refiner_and_condition_tokens_vs_oracle():
    """Refiner on compacted valid tokens vs the reference's masked refiner (valid rows), and the valid-first token order."""
    name = "hy15_t2v"
    cfg, g = CONFIGS[name], load(name)
    w32 = hy15_dit.make_weights(**cfg, seed=1234, dtype=torch.float32)
    m = _model(cfg, w32)
    x, t, text, mask, text2, mask2, img = inputs(g, torch.bfloat16)
    valid = mask[0].bool()
    got = m.token_refiner(text[0][valid].to(DEV).contiguous(), t.to(DEV))
    exact = torch.from_numpy(g["refined_fp32"])[0][valid]
    ref16 = torch.from_numpy(g["refined_bf16"])[0][valid]
    assert rel_l2(got, exact) <= max(1e-3, 1.5 * rel_l2(ref16, exact)), (rel_l2(got, exact), rel_l2(ref16, exact))
    ctx = m.condition_tokens(text[0].to(DEV), mask[0], text2[0].to(DEV), mask2[0], img[0].to(DEV), t.to(DEV))
    x32, t32, text32, _, text2_32, _, img32 = inputs(g, torch.float32)
    want = hy15_dit.condition_tokens(text32, mask, text2_32, mask2, img32, t32, w32, cfg["heads"], cfg["num_refiner_layers"])[0]
    assert ctx.shape == want.shape
    n_valid = int(mask2.sum() + mask.sum())
    n_img = img.shape[1]
    assert rel_l2(ctx[:n_valid], want[:n_valid]) <= 1e-2
    assert torch.equal(ctx[n_valid:n_valid + n_img].float().cpu(), w32["cond_type_embed.weight"][2].bfloat16().float().expand(n_img, -1))
    assert ctx[n_valid + n_img:].abs().max().item() == 0          # padded text tokens are zeros (model.py:1073-1074)


def test_hy15_full_width_one_block_vs_exact_oracle():
    """HunyuanVideo-1.5 widths (d = 2048, 16 x 128 heads, text 3584 / ByT5 1472 / image 1152, 65 latent channels) with one
    dual-stream block and one refiner block on a 5 x 16 x 20 latent grid (1600 tokens) + 64 + 32 + 16 condition tokens."""
    cfg = dict(dim=2048, heads=16, num_layers=1, num_refiner_layers=1, in_channels=65, out_channels=32, text_dim=3584,
               text2_dim=1472, image_dim=1152, byt5_hidden=2048)
    w32 = hy15_dit.make_weights(**cfg, seed=5, dtype=torch.float32, std=0.02)
    from apex_studio_b200.hunyuanvideo15 import HunyuanVideo15Config, HunyuanVideo15Transformer3DModel

    m = HunyuanVideo15Transformer3DModel(HunyuanVideo15Config(num_layers=1, num_refiner_layers=1))
    m.load_state_dict(w32, device=DEV)
    gen = torch.Generator().manual_seed(42)
    x = torch.randn(1, 65, 5, 16, 20, generator=gen)
    text, text2, img = torch.randn(1, 64, 3584, generator=gen), torch.randn(1, 32, 1472, generator=gen), torch.randn(1, 16, 1152, generator=gen)
    mask, mask2 = torch.ones(1, 64), torch.ones(1, 32)
    mask[:, 50:], mask2[:, 20:] = 0, 0
    t = torch.tensor([750.0])
    out = _call(m, x.bfloat16(), t.bfloat16(), text.bfloat16(), mask, text2.bfloat16(), mask2, img.bfloat16())
    k = dict(heads=16, num_layers=1, num_refiner_layers=1)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    exact = hy15_dit.hy15_forward(x, t, text, mask, text2, mask2, img, w32, **k)
    bf = hy15_dit.hy15_forward(x.bfloat16(), t.bfloat16(), text.bfloat16(), mask, text2.bfloat16(), mask2, img.bfloat16(),
                               {kk: v.bfloat16() for kk, v in w32.items()}, **k)
    ours, theirs = rel_l2(out, exact), rel_l2(bf, exact)
    assert torch.isfinite(out).all() and tuple(out.shape) == (1, 32, 5, 16, 20)
    assert ours <= max(1e-3, 1.5 * theirs), (ours, theirs)


def test_hy15_condition_plan_is_cached_per_prompt_and_invalidated():
    """The timestep-independent part of condition_tokens (valid counts, t2v flag, compacted tokens, image / byT5 branches) is
    computed once per prompt: the same device tensors hit the cached plan (no host syncs, same result), an in-place change
    of a mask or new weights miss it."""
    name = "hy15_t2v"
    cfg, g = CONFIGS[name], load(name)
    w32 = hy15_dit.make_weights(**cfg, seed=1234, dtype=torch.float32)
    m = _model(cfg, w32)
    x, t, text, mask, text2, mask2, img = inputs(g, torch.bfloat16)
    dev_args = [text[0].to(DEV), mask[0].to(DEV), text2[0].to(DEV), mask2[0].to(DEV), img[0].to(DEV)]
    a = m.condition_tokens(*dev_args, t.to(DEV))
    assert len(m._cond_plans) == 1
    b = m.condition_tokens(*dev_args, t.to(DEV))
    assert len(m._cond_plans) == 1 and torch.equal(a, b)
    c = m.condition_tokens(*dev_args, (t * 0.5).to(DEV))            # only the token refiner depends on the timestep
    assert len(m._cond_plans) == 1 and not torch.equal(a, c)
    n_valid = int(dev_args[1].sum())
    dev_args[1][n_valid - 1] = 0                                    # in-place edit of the mask: a new plan, one token fewer
    d = m.condition_tokens(*dev_args, t.to(DEV))
    assert len(m._cond_plans) == 2 and d.shape == a.shape
    assert d[-1].abs().max().item() == 0 and not torch.equal(a, d)
    m.load_state_dict(w32, device=DEV)
    assert len(m._cond_plans) == 0


def test_hy15_forward_cuda_graph_replay_is_bit_identical_to_eager():
    """With the per-prompt condition plan the forward has no host <-> device synchronisation left: one whole forward is captured
    into a CUDA graph (latents and timestep are graph inputs, the prompt tensors are bound at capture) and replayed with new
    latents / timesteps bit-identically to eager execution (as the Flux and Wan forwards already are)."""
    from apex_studio_b200.graph import GraphedCallable

    name = "hy15_t2v"
    cfg, g = CONFIGS[name], load(name)
    m = _model(cfg, hy15_dit.make_weights(**cfg, seed=1234, dtype=torch.float32))
    x, t, text, mask, text2, mask2, img = inputs(g, torch.bfloat16)
    x, t = x.to(DEV), t.to(DEV, torch.bfloat16)
    prompt = dict(encoder_hidden_states=text.to(DEV), encoder_attention_mask=mask.to(DEV), encoder_hidden_states_2=text2.to(DEV),
                  encoder_attention_mask_2=mask2.to(DEV), image_embeds=img.to(DEV))
    fwd = lambda xx, tt: m(xx, tt, return_dict=False, **prompt)[0]
    graphed = GraphedCallable(fwd, (x, t))
    x2, t2 = (x.float() * 0.5 + 0.1).bfloat16(), (t.float() * 0.25).bfloat16()
    for xi, ti in ((x, t), (x2, t2), (x, t2)):
        eager = fwd(xi, ti).clone()
        assert torch.equal(graphed(xi, ti), eager)
