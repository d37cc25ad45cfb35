"""Multi-process host logic on CPU (gloo, world_size 2 and 4): CFG-pair exchange, Ulysses token<->head
re-partitioning, the CFG-parallel denoise loop.  The arithmetic kernels are replaced by the oracle here
(this file tests the HOST side of the sharding; kernels are tested in test_gpu_parity.py)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, fn, *args):
    port = _free_port()
    mp.spawn(_entry, args=(world, port, fn, args), nprocs=world, join=True)


def _entry(rank, world, port, fn, args):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
def _ulysses_roundtrip(rank, world, use_cfg):
    from apex_studio_b200.parallel import ParallelContext

    par = ParallelContext.create(use_cfg=use_cfg)
    P = par.sp_size
    heads, hd, S = 4, 8, 24
    g = torch.Generator().manual_seed(5)
    qkv_full = torch.randn(S, 3 * heads * hd, generator=g)       # same on every rank
    lo, hi = par.shard_bounds(S)
    got = par.tokens_to_heads(qkv_full[lo:hi].contiguous(), heads, hd)   # [3, S, (H/P)*hd]
    hp = heads // P
    for which in range(3):
        cols = qkv_full[:, which * heads * hd:(which + 1) * heads * hd]
        expect = cols[:, par.sp_rank * hp * hd:(par.sp_rank + 1) * hp * hd]
        assert torch.equal(got[which], expect), (rank, which)
    # "attention output" for my heads = q for my heads; after heads_to_tokens I must hold q of my tokens
    back = par.heads_to_tokens(got[0].contiguous())
    assert torch.equal(back, qkv_full[lo:hi, :heads * hd]), rank
    full = par.gather_tokens(back)
    assert torch.equal(full, qkv_full[:, :heads * hd])


@pytest.mark.parametrize("world,use_cfg", [(2, False), (4, True), (4, False)])
def test_ulysses_repartition(world, use_cfg):
    _run(world, _ulysses_roundtrip, use_cfg)


def _cfg_loop(rank, world):
    """CFG-parallel moe_denoise on 2 ranks == sequential cond/uncond loop on one rank (bit for bit)."""
    import numpy as np

    import wan_dit
    from apex_studio_b200 import denoise, ops
    from apex_studio_b200.parallel import ParallelContext
    from apex_studio_b200.scheduler import UniPCMultistepScheduler

    ops.cfg_combine = lambda c, u, g: wan_dit.cfg_combine(c, u, g)   # oracle arithmetic on CPU

    class FakeDiT:
        def __init__(self, k):
            self.k = k

        def __call__(self, hidden_states, timestep, encoder_hidden_states, return_dict=False, parallel=None, **kw):
            e = encoder_hidden_states.float().mean()
            y = self.k * hidden_states.float() + 0.05 * torch.sin(hidden_states.float() * 2 + e + timestep.float().view(-1, 1, 1, 1, 1) * 1e-3)
            return (y.to(hidden_states.dtype),)

    def run(par):
        sch = UniPCMultistepScheduler(shift=3.0)
        sch.set_timesteps(6)
        lat = torch.randn(1, 4, 2, 4, 4, generator=torch.Generator().manual_seed(1))
        tr = denoise.DenoiseTrace()
        out = denoise.moe_denoise(
            timesteps=sch.timesteps, latents=lat, scheduler=sch, high_noise_transformer=FakeDiT(0.3),
            low_noise_transformer=FakeDiT(0.4), boundary_timestep=875.0, guidance_scale=[4.0, 3.0],
            transformer_kwargs=dict(encoder_hidden_states=torch.ones(1, 3, 8)),
            unconditional_transformer_kwargs=dict(encoder_hidden_states=torch.zeros(1, 3, 8)), parallel=par, trace=tr)
        return out, tr

    par = ParallelContext.create(use_cfg=True)
    assert (par.cfg_size, par.sp_size) == (2, 1)
    sharded, tr_p = run(par)
    single, tr_s = run(ParallelContext.single())
    assert torch.equal(sharded, single)
    assert tr_p.timesteps == tr_s.timesteps and tr_p.expert == tr_s.expert and tr_p.guidance == tr_s.guidance
    assert tr_s.expert[0] == "high" and tr_s.expert[-1] == "low"
    # both ranks hold identical latents without any broadcast
    both = [torch.empty_like(sharded) for _ in range(world)]
    dist.all_gather(both, sharded)
    assert torch.equal(both[0], both[1])


def test_cfg_parallel_denoise_loop():
    _run(2, _cfg_loop)


def test_plan_layout_and_dealing():
    from apex_studio_b200.parallel import deal_round_robin, plan_layout

    assert plan_layout(1, True) == (1, 1) and plan_layout(2, True) == (2, 1)
    assert plan_layout(4, True) == (2, 2) and plan_layout(8, True) == (2, 4) and plan_layout(8, False) == (1, 8)
    tiles = [deal_round_robin(28, 8, r) for r in range(8)]
    assert sorted(sum(tiles, [])) == list(range(28)) and [len(t) for t in tiles] == [4, 4, 4, 4, 3, 3, 3, 3]
    # cost-aware dealing of the 28 Wan VAE tiles of a 90 x 160 latent (tile 32, stride 24): the heaviest of 8 ranks carries
    # 3136 of 23,712 latent pixels (7.56 x), round-robin gives 3648 (6.5 x)
    from apex_studio_b200.parallel import deal_lpt

    grid = [(i, j) for i in range(0, 90, 24) for j in range(0, 160, 24)]
    costs = [min(32, 90 - i) * min(32, 160 - j) for i, j in grid]
    for world in (1, 2, 4, 8):
        owners = deal_lpt(costs, world)
        assert sorted(sum(owners, [])) == list(range(28)) and owners == deal_lpt(costs, world)
        loads = [sum(costs[i] for i in o) for o in owners]
        assert max(loads) <= sum(costs) / world * 1.06
    loads8 = [sum(costs[i] for i in o) for o in deal_lpt(costs, 8)]
    rr8 = [sum(costs[i] for i in deal_round_robin(28, 8, r)) for r in range(8)]
    assert max(loads8) == 3136 and sum(costs) / max(loads8) > 7.5 and sum(costs) / max(rr8) <= 6.5


# ---------------------------------------------------------------------------------------------------
def _joint_exchange(rank, world, img_first):
    """Dual-stream families: image stream token-sharded, text stream replicated (ParallelContext.joint_*)."""
    from apex_studio_b200.parallel import ParallelContext

    par = ParallelContext.create(use_cfg=False)
    P = par.sp_size
    heads, hd, n_img, n_txt = 4, 8, 24, 5
    d = heads * hd
    g = torch.Generator().manual_seed(9)
    img_full = torch.randn(n_img, 3 * d, generator=g)           # same on every rank
    txt_full = torch.randn(n_txt, 3 * d, generator=g)
    lo, hi = par.shard_bounds(n_img)
    n_loc = hi - lo
    if img_first:
        local = torch.cat([img_full[lo:hi], txt_full], dim=0)
        img_rows, txt_rows = slice(0, n_loc), slice(n_loc, None)
        joint = torch.cat([img_full, txt_full], dim=0)
    else:
        local = torch.cat([txt_full, img_full[lo:hi]], dim=0)
        img_rows, txt_rows = slice(n_txt, None), slice(0, n_txt)
        joint = torch.cat([txt_full, img_full], dim=0)
    got = par.joint_tokens_to_heads(local, img_rows, txt_rows, heads, hd, img_first)       # [3, S_joint, (H/P)*hd]
    hp = heads // P
    for which in range(3):
        cols = joint[:, which * d:(which + 1) * d]
        assert torch.equal(got[which], cols[:, par.sp_rank * hp * hd:(par.sp_rank + 1) * hp * hd]), (rank, which)
    # "attention output" for my heads = q of my heads: afterwards I must hold q (all heads) of my image rows + all text rows
    out = torch.zeros(n_loc + n_txt, d)
    par.joint_heads_to_tokens(got[0].contiguous(), n_img, img_first, out, img_rows, txt_rows)
    assert torch.equal(out[img_rows], img_full[lo:hi, :d]) and torch.equal(out[txt_rows], txt_full[:, :d]), rank


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("img_first", [True, False])
def test_dual_stream_joint_exchange(world, img_first):
    _run(world, _joint_exchange, img_first)


def _hy15_cfg_loop(rank, world):
    """CFG-parallel hy15_denoise on 2 ranks == the sequential uncond/cond loop of engine/hunyuanvideo15/t2v.py:234-338."""
    import numpy as np

    import wan_dit
    from apex_studio_b200 import denoise, ops
    from apex_studio_b200.parallel import ParallelContext
    from apex_studio_b200.scheduler import FlowMatchEulerDiscreteScheduler

    ops.cfg_combine = lambda c, u, g: wan_dit.cfg_combine(c, u, g)   # oracle arithmetic on CPU
    calls = []

    class FakeDiT:
        def __call__(self, hidden_states, timestep, encoder_hidden_states, image_embeds, return_dict=False, parallel=None, **kw):
            assert hidden_states.shape[1] == 9 and timestep.dtype == hidden_states.dtype           # cat(latents, cond, mask)
            calls.append(float(encoder_hidden_states.float().mean()))
            e = encoder_hidden_states.float().mean()
            y = 0.3 * hidden_states[:, :4].float() + 0.05 * torch.sin(hidden_states[:, :4].float() * 2 + e + timestep.float().view(-1, 1, 1, 1, 1) * 1e-3)
            return (y.to(hidden_states.dtype),)

    def run(par):
        sch = FlowMatchEulerDiscreteScheduler(use_dynamic_shifting=False, shift=7.0)
        ts = sch.set_timesteps(5, sigmas=np.linspace(1.0, 0.0, 6)[:-1])
        lat = torch.randn(1, 4, 2, 4, 4, generator=torch.Generator().manual_seed(1)).bfloat16()
        kw = dict(encoder_attention_mask=torch.ones(1, 3), encoder_hidden_states_2=torch.zeros(1, 2, 8), encoder_attention_mask_2=torch.ones(1, 2))
        return denoise.hy15_denoise(timesteps=ts, latents=lat, scheduler=sch, transformer=FakeDiT(),
                                    cond_latents_concat=torch.zeros(1, 4, 2, 4, 4).bfloat16(), mask_concat=torch.zeros(1, 1, 2, 4, 4).bfloat16(),
                                    image_embeds=torch.zeros(1, 2, 8), cond_kwargs=dict(kw, encoder_hidden_states=torch.ones(1, 3, 8)),
                                    uncond_kwargs=dict(kw, encoder_hidden_states=torch.zeros(1, 3, 8)), guidance_scale=6.0, parallel=par)

    par = ParallelContext.create(use_cfg=True)
    calls.clear()
    sharded = run(par)
    assert len(calls) == 5 and set(calls) == ({1.0} if par.cfg_rank == 0 else {0.0})      # one branch per rank
    calls.clear()
    single = run(ParallelContext.single())
    assert calls[:2] == [0.0, 1.0] and len(calls) == 10                                   # uncond first, then cond (t2v.py:268-289)
    assert sharded.dtype == torch.bfloat16 and torch.equal(sharded, single)
    with pytest.raises(ValueError):
        denoise.hy15_denoise(timesteps=[], latents=None, scheduler=None, transformer=None, cond_latents_concat=None, mask_concat=None,
                             image_embeds=None, cond_kwargs={}, uncond_kwargs=None, do_classifier_free_guidance=True)


def test_hy15_cfg_parallel_denoise_loop():
    _run(2, _hy15_cfg_loop)


def test_flux_denoise_loop_host_logic():
    """flux_denoise (engine/flux/shared.py:504-620) with a fake transformer on the CPU: timestep / 1000 in the latent dtype,
    guidance passed through, true-CFG branch order and combine, scheduler stepped once per timestep."""
    import numpy as np

    import wan_dit
    from apex_studio_b200 import denoise, ops
    from apex_studio_b200.scheduler import FlowMatchEulerDiscreteScheduler, calculate_shift

    ops.cfg_combine = lambda c, u, g: wan_dit.cfg_combine(c, u, g)
    seen = []

    def fake(hidden_states, timestep, guidance, img_ids, pooled_projections, encoder_hidden_states, txt_ids, return_dict=False):
        seen.append((float(timestep[0]), float(pooled_projections.mean())))
        assert timestep.dtype == hidden_states.dtype and float(timestep[0]) <= 1.0 and float(guidance[0]) == 3.5
        return (0.5 * hidden_states + pooled_projections.mean().to(hidden_states.dtype),)

    sch = FlowMatchEulerDiscreteScheduler()
    ts = sch.set_timesteps(4, sigmas=np.linspace(1.0, 0.25, 4), mu=calculate_shift(16))
    lat = torch.randn(1, 16, 8, generator=torch.Generator().manual_seed(0)).bfloat16()
    out = denoise.flux_denoise(latents=lat, timesteps=ts, scheduler=sch, transformer=fake, prompt_embeds=torch.zeros(1, 2, 4),
                               pooled_prompt_embeds=torch.ones(1, 4), latent_ids=torch.zeros(16, 3), text_ids=torch.zeros(2, 3),
                               guidance=torch.full([1], 3.5), negative_prompt_embeds=torch.zeros(1, 2, 4),
                               negative_pooled_prompt_embeds=torch.zeros(1, 4), negative_text_ids=torch.zeros(2, 3),
                               true_cfg_scale=2.0, use_cfg_guidance=True)
    assert out.shape == lat.shape and out.dtype == torch.bfloat16 and sch.step_index == 4
    assert len(seen) == 8 and [p for _, p in seen[:2]] == [1.0, 0.0]                    # positive branch first, then negative
    assert seen[0][0] == pytest.approx(1.0, abs=1e-2)
