"""Multi-GPU correctness check of the dual-stream families (run under torchrun on 2 / 4 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 \
        scripts/gpu_mp_check_mmdit.py

On every rank: (1) sequence-parallel HunyuanVideo-1.5 forward == single-GPU forward (NCCL exchange and the exchange fused into
the kernels over NVLink peer memory), (2) the same for QwenImage (edit, two
images), (3) CFG x SP hy15_denoise == the sequential loop, (4) tile-parallel HunyuanVideo-1.5 VAE decode == single-GPU
tiled decode.  Prints one JSON line from rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import hy15_dit
    import hy15_vae
    import qwen_dit
    from apex_studio_b200 import denoise
    from apex_studio_b200.hunyuanvideo15 import HunyuanVideo15Config, HunyuanVideo15Transformer3DModel
    from apex_studio_b200.parallel import ParallelContext
    from apex_studio_b200.qwenimage import QwenImageConfig, QwenImageTransformer2DModel
    from apex_studio_b200.scheduler import FlowMatchEulerDiscreteScheduler
    from apex_studio_b200.vae import AutoencoderKLHunyuanVideo15, HunyuanVideo15VAEConfig

    res = {"world": world}
    bf = torch.bfloat16
    par_sp = ParallelContext.create(use_cfg=False)
    par = ParallelContext.create(use_cfg=True)
    res["sp_layout"], res["cfg_layout"] = [par_sp.cfg_size, par_sp.sp_size], [par.cfg_size, par.sp_size]

    # 1. HunyuanVideo-1.5, 4 heads, 256 latent tokens + 21 condition tokens
    cfg = dict(dim=512, heads=4, num_layers=2, num_refiner_layers=1, in_channels=9, out_channels=4, text_dim=48, text2_dim=40,
               image_dim=24, byt5_hidden=64)
    m = HunyuanVideo15Transformer3DModel(HunyuanVideo15Config(in_channels=9, out_channels=4, num_attention_heads=4, num_layers=2,
                                                              num_refiner_layers=1, text_embed_dim=48, text_embed_2_dim=40,
                                                              image_embed_dim=24))
    m.load_state_dict(hy15_dit.make_weights(**cfg, seed=3), device=dev)
    g = torch.Generator().manual_seed(42)
    x = torch.randn(1, 9, 4, 8, 8, generator=g).to(dev, bf)
    text, text2 = torch.randn(1, 10, 48, generator=g).to(dev, bf), torch.randn(1, 6, 40, generator=g).to(dev, bf)
    neg = torch.randn(1, 10, 48, generator=g).to(dev, bf)
    img = torch.zeros(1, 5, 24, device=dev, dtype=bf)
    m1, m2 = torch.ones(1, 10), torch.ones(1, 6)
    m1[:, 7:], m2[:, 4:] = 0, 0
    t = torch.tensor([500.0], device=dev, dtype=bf)
    kw = dict(encoder_hidden_states=text, encoder_attention_mask=m1, encoder_hidden_states_2=text2, encoder_attention_mask_2=m2)
    single = m(x, t, image_embeds=img, return_dict=False, **kw)[0]
    sharded = m(x, t, image_embeds=img, return_dict=False, parallel=par_sp, **kw)[0]
    res["hy15_sp_max_abs_diff"] = (single.float() - sharded.float()).abs().max().item()
    # the same with the exchange fused into the kernels over NVLink peer memory (parallel.JointPeerExchange), twice (buffer reuse)
    par_p2p = ParallelContext.create(use_cfg=False, use_p2p=True)
    for rep in range(2):
        fused = m(x, t, image_embeds=img, return_dict=False, parallel=par_p2p, **kw)[0]
        res[f"hy15_sp_p2p_run{rep}_max_abs_diff"] = (single.float() - fused.float()).abs().max().item()

    # 2. QwenImage edit: 64 + 32 image tokens, 13 text tokens
    qcfg = dict(dim=512, heads=4, num_layers=2, in_channels=16, out_channels=4, joint_dim=48)
    q = QwenImageTransformer2DModel(QwenImageConfig(in_channels=16, out_channels=4, num_layers=2, num_attention_heads=4,
                                                    joint_attention_dim=48))
    q.load_state_dict(qwen_dit.make_weights(**qcfg, seed=4), device=dev)
    shapes = [(1, 8, 8), (1, 4, 8)]
    qx, qe = torch.randn(1, 96, 16, generator=g).to(dev, bf), torch.randn(1, 13, 48, generator=g).to(dev, bf)
    qt = torch.tensor([0.5], device=dev)
    qkw = dict(hidden_states=qx, encoder_hidden_states=qe, timestep=qt, img_shapes=[shapes], txt_seq_lens=[13], return_dict=False)
    q_single = q(**qkw)[0]
    res["qwen_sp_max_abs_diff"] = (q_single.float() - q(parallel=par_sp, **qkw)[0].float()).abs().max().item()
    for rep in range(2):   # text rows FIRST in QwenImage's joint sequence: the replicated rows lead
        res[f"qwen_sp_p2p_run{rep}_max_abs_diff"] = (q_single.float() - q(parallel=par_p2p, **qkw)[0].float()).abs().max().item()

    # 3. CFG x SP denoise loop (HunyuanVideo-1.5: latents 4 ch + cond 4 ch + mask 1 ch = 9 input channels)
    lat = torch.randn(1, 4, 4, 8, 8, generator=g).to(dev, bf)
    cond_lat, mask = torch.zeros(1, 4, 4, 8, 8, device=dev, dtype=bf), torch.zeros(1, 1, 4, 8, 8, device=dev, dtype=bf)

    def run(p):
        sch = FlowMatchEulerDiscreteScheduler(use_dynamic_shifting=False, shift=7.0)
        ts = sch.set_timesteps(3, device=dev, sigmas=np.linspace(1.0, 0.0, 4)[:-1])
        return denoise.hy15_denoise(timesteps=ts, latents=lat.clone(), scheduler=sch, transformer=m, cond_latents_concat=cond_lat,
                                    mask_concat=mask, image_embeds=img, cond_kwargs=kw, uncond_kwargs=dict(kw, encoder_hidden_states=neg),
                                    guidance_scale=6.0, parallel=p)

    a, b = run(par), run(ParallelContext.single())
    res["denoise_max_abs_diff"] = (a.float() - b.float()).abs().max().item()
    res["denoise_finite"] = bool(torch.isfinite(a).all())

    # 4. tile-parallel VAE decode
    ch = (128, 128, 64, 64, 32)
    vae = AutoencoderKLHunyuanVideo15(HunyuanVideo15VAEConfig(block_out_channels=tuple(reversed(ch))))
    vae.load_state_dict(hy15_vae.make_weights(ch, seed=7), device=dev)
    vae.enable_tiling()
    z = torch.randn(1, 32, 2, 14, 16, generator=torch.Generator().manual_seed(5)).to(dev, bf)
    par_all = ParallelContext.create(use_cfg=False)
    v1 = vae.decode(z, return_dict=False)[0]
    v2 = vae.decode(z, return_dict=False, parallel=par_all)[0]
    res["vae_tiles"] = len(vae.tile_grid(14, 16))
    res["vae_max_abs_diff"] = (v1.float() - v2.float()).abs().max().item()

    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        print(json.dumps({"per_rank": gathered}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
