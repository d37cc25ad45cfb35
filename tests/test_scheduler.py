"""Scheduler: integer step indexing bit-exact vs the reference (golden from the reference class itself),
float update within 2e-6 (oracle, numpy) / exact (product, torch on CPU)."""
import os

import numpy as np
import pytest
import torch

import unipc
from apex_studio_b200.denoise import select_expert_is_high, select_guidance_scale
from apex_studio_b200.scheduler import UniPCMultistepScheduler, get_timesteps
from conftest import GOLDEN

CASES = [(4, 3.0), (8, 5.0), (50, 3.0), (50, 1.0)]


def synthetic_model(sample, t):
    return 0.35 * sample + 0.1 * torch.sin(sample * 3.0 + float(t) * 0.01)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "unipc.npz"))


@pytest.mark.parametrize("n,shift", CASES)
def test_integer_schedule_bit_exact(gold, n, shift):
    tag = f"n{n}_s{shift:g}"
    sig, ts = unipc.make_schedule(n, shift)
    assert ts.dtype == np.int64 and np.array_equal(ts, gold[tag + "_timesteps"])
    assert np.array_equal(sig, gold[tag + "_sigmas"])
    s = UniPCMultistepScheduler(shift=shift)
    s.set_timesteps(n)
    assert s.timesteps.dtype == torch.int64 and np.array_equal(s.timesteps.numpy(), gold[tag + "_timesteps"])
    assert np.array_equal(s.sigmas.numpy(), gold[tag + "_sigmas"])
    assert np.array_equal(np.array([(a, b, int(c)) for a, b, c in unipc.step_orders(n)]), gold[tag + "_trace"])


@pytest.mark.parametrize("n,shift", CASES)
def test_full_run_vs_reference(gold, n, shift):
    tag = f"n{n}_s{shift:g}"
    ref = gold[tag + "_final"]
    # product scheduler (torch, CPU here / GPU in production)
    s = UniPCMultistepScheduler(shift=shift)
    s.set_timesteps(n)
    x = torch.from_numpy(gold[tag + "_x0"].copy())
    norms = []
    for t in s.timesteps:
        x = s.step(synthetic_model(x, int(t)), t, x)[0]
        norms.append(float(x.double().norm()))
    assert np.array_equal(np.array([(a, b, int(c)) for a, b, c in s.trace]), gold[tag + "_trace"])
    assert np.abs(x.numpy() - ref).max() <= 2e-6 * np.abs(ref).max()
    assert np.allclose(norms, gold[tag + "_norms"], rtol=2e-6)
    # numpy oracle
    sig, ts = unipc.make_schedule(n, shift)
    o = unipc.UniPCOracle(sig, ts)
    y = gold[tag + "_x0"].copy()
    for t in ts:
        y = o.step(synthetic_model(torch.from_numpy(y), int(t)).numpy(), int(t), y)
    assert np.array_equal(np.array([(a, b, int(c)) for a, b, c in o.trace]), gold[tag + "_trace"])
    assert np.abs(y - ref).max() <= 2e-6 * np.abs(ref).max()


def test_bf16_model_output_promotes_like_the_reference(gold):
    """noise_pred is bf16, latents fp32: sigma * model_output is a bf16 product (0-dim tensors do not promote)."""
    s = UniPCMultistepScheduler(shift=3.0)
    s.set_timesteps(4)
    x = torch.from_numpy(gold["n4_s3_x0"].copy())
    mo = synthetic_model(x, 999).bfloat16()
    s._step_index = 0
    conv = s.convert_model_output(mo, x)
    assert conv.dtype == torch.float32
    assert torch.equal(conv, x - (s.sigmas[0] * mo))
    assert (s.sigmas[0] * mo).dtype == torch.bfloat16


def test_index_for_timestep_second_match_rule():
    s = UniPCMultistepScheduler()
    s.set_timesteps(4)
    s.timesteps = torch.tensor([900, 900, 500, 100])
    assert s.index_for_timestep(900) == 1 and s.index_for_timestep(torch.tensor(500)) == 2
    assert unipc.index_for_timestep(np.array([900, 900, 500, 100]), 900) == 1


def test_disable_corrector_and_lower_order_final():
    tr = unipc.step_orders(6, solver_order=3, disable_corrector=[0, 2])
    assert [o for _, o, _ in tr] == [1, 2, 3, 3, 2, 1]
    assert [c for _, _, c in tr] == [False, False, True, False, True, True]
    s = UniPCMultistepScheduler(solver_order=3, disable_corrector=[0, 2])
    s.set_timesteps(6)
    x = torch.randn(2, 3)
    for t in s.timesteps:
        x = s.step(synthetic_model(x, int(t)), t, x)[0]
    assert s.trace == tr and torch.isfinite(x).all()


def test_get_timesteps_variants():
    s = UniPCMultistepScheduler(shift=3.0)
    ts, n = get_timesteps(s, num_inference_steps=10)
    assert n == 10 and np.array_equal(ts.numpy(), unipc.make_schedule(10, 3.0)[1])
    ts2, n2 = get_timesteps(s, num_inference_steps=10, strength=0.5)
    assert n2 == 5 and np.array_equal(ts2.numpy(), unipc.strength_cut(ts.numpy(), 0.5))
    s.set_timesteps(1000)
    full = s.timesteps.clone()
    ids = [1000, 750, 500, 250]
    ts3, n3 = get_timesteps(s, timesteps=ids, timesteps_as_indices=True)
    assert n3 == 4 and np.array_equal(ts3.numpy(), unipc.timesteps_as_indices(full.numpy(), ids))


def test_expert_switch_is_integer_exact():
    _, ts = unipc.make_schedule(50, 3.0)
    boundary = 0.875 * 1000
    expect = unipc.expert_and_guidance(ts, boundary, [4.0, 3.0])
    got = [("high" if select_expert_is_high(torch.tensor(t), boundary) else "low",
            select_guidance_scale(torch.tensor(t), boundary, [4.0, 3.0])) for t in ts]
    assert got == expect
    assert sum(e == "high" for e, _ in got) == int((ts >= 875).sum())
    assert select_guidance_scale(torch.tensor(10), None, 5.0) == 5.0


def test_preview_renderer_follows_the_reference_render_condition():
    """render_on_step (engine/wan/shared/__init__.py:580-586): first step, every `interval`-th step, never the last one."""
    import numpy as np

    from apex_studio_b200 import denoise

    got = []
    pr = denoise.PreviewRenderer(lambda lat: (lat[0, :3, 0].permute(1, 2, 0).abs().clamp(0, 1) * 255).to(torch.uint8)[None],
                                 lambda frames: got.append(frames), interval=3)
    total = 10
    for i in range(total):
        pr.maybe_render(i, total, torch.full((1, 4, 2, 3, 5), float(i) / 10))
    pr.finish()
    assert pr.rendered_steps == [0, 2, 5, 8] and len(got) == 4
    assert all(isinstance(f, np.ndarray) and f.dtype == np.uint8 and f.shape == (1, 3, 5, 3) for f in got)
    assert got[1][0, 0, 0, 0] == int(0.2 * 255)
    with pytest.raises(ValueError):
        denoise.PreviewRenderer(lambda x: x, lambda f: None, interval=0)

    class Sch:
        def step(self, mo, t, x, return_dict=False):
            return (x - 0.1 * mo.float(),)

    fake = lambda hidden_states, timestep, return_dict=False, parallel=None, **kw: (hidden_states * 0.5,)
    frames = []
    out = denoise.moe_denoise(timesteps=torch.arange(5, 0, -1), latents=torch.ones(1, 4, 1, 2, 2), scheduler=Sch(),
                              high_noise_transformer=fake, use_cfg_guidance=False, render_on_step=True,
                              render_on_step_callback=frames.append, render_on_step_interval=2,
                              preview_decode_fn=lambda lat: (lat[0, :3, 0].permute(1, 2, 0).clamp(0, 1) * 255).to(torch.uint8)[None])
    assert len(frames) == 3 and out.shape == (1, 4, 1, 2, 2)          # steps 0, 1, 3 of 5 (never the last)
    with pytest.raises(ValueError):
        denoise.moe_denoise(timesteps=torch.arange(2), latents=torch.ones(1, 4, 1, 2, 2), scheduler=Sch(), high_noise_transformer=fake,
                            use_cfg_guidance=False, render_on_step=True, render_on_step_callback=frames.append)


def test_step_returns_scheduler_output_by_default_and_euler_starts_at_its_timestep():
    """(1) ``step`` defaults to return_dict=True and returns an object with ``.prev_sample`` (reference
    scheduler/unipc.py:651-737) that is also indexable; (2) FlowMatchEuler without ``set_begin_index`` starts at the index
    of the timestep it is given (upstream ``_init_step_index``), e.g. on a strength-truncated schedule."""
    from apex_studio_b200.scheduler import FlowMatchEulerDiscreteScheduler, SchedulerOutput, UniPCMultistepScheduler

    s = UniPCMultistepScheduler(shift=3.0)
    s.set_timesteps(4)
    x = torch.randn(1, 4, 2, 2, 2)
    out = s.step(torch.zeros_like(x), s.timesteps[0], x)
    assert isinstance(out, SchedulerOutput) and torch.equal(out.prev_sample, out[0])
    s2 = UniPCMultistepScheduler(shift=3.0)
    s2.set_timesteps(4)
    assert isinstance(s2.step(torch.zeros_like(x), s2.timesteps[0], x, return_dict=False), tuple)

    e = FlowMatchEulerDiscreteScheduler(shift=3.0, use_dynamic_shifting=False)
    e.set_timesteps(10)
    mo = torch.ones_like(x)
    start = 4                                        # img2img: run only timesteps[4:]
    got = e.step(mo, e.timesteps[start], x).prev_sample
    want = x + (e.sigmas[start + 1] - e.sigmas[start]) * mo
    assert torch.allclose(got, want) and e._step_index == start + 1
    e2 = FlowMatchEulerDiscreteScheduler(shift=3.0, use_dynamic_shifting=False)
    e2.set_timesteps(10)
    e2.set_begin_index(2)
    assert torch.allclose(e2.step(mo, e2.timesteps[2], x)[0], x + (e2.sigmas[3] - e2.sigmas[2]) * mo)
