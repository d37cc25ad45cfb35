"""Multi-GPU correctness check (run under torchrun on 2 / 4 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/gpu_mp_check.py

Checks, on every rank, that the sharded paths give the SAME result as the unsharded path computed locally:
  1. sequence-parallel DiT forward (Ulysses token<->head all-to-all) == single-GPU forward
  2. CFG-parallel + sequence-parallel moe_denoise == sequential loop
  3. tile-parallel VAE decode (round-robin tiles + one all-gather) == single-GPU tiled decode
Prints one JSON line from rank 0.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import wan_dit
    import wan_vae
    from apex_studio_b200 import denoise
    from apex_studio_b200.parallel import ParallelContext
    from apex_studio_b200.scheduler import UniPCMultistepScheduler
    from apex_studio_b200.vae import AutoencoderKLWan, WanVAEConfig
    from apex_studio_b200.wan import WanConfig, WanTransformer3DModel

    res = {"world": world}
    cfg = dict(dim=1024, heads=8, ffn_dim=1536, num_layers=2, text_dim=64, freq_dim=256)
    w32 = wan_dit.make_weights(**cfg, seed=1234)
    model = WanTransformer3DModel(WanConfig(num_attention_heads=8, text_dim=64, ffn_dim=1536, num_layers=2))
    model.load_state_dict(w32, device=dev)
    g = torch.Generator().manual_seed(42)
    lat = torch.randn(1, 16, 4, 16, 24, generator=g)          # 4 x 8 x 12 = 384 tokens
    text = torch.randn(1, 24, 64, generator=g).bfloat16()
    neg = torch.randn(1, 24, 64, generator=g).bfloat16()
    t = torch.tensor([900], device=dev)

    # 1. sequence parallel over ALL ranks
    par_sp = ParallelContext.create(use_cfg=False)
    single = model(lat.to(dev, torch.bfloat16), t, text.to(dev), return_dict=False)[0]
    sharded = model(lat.to(dev, torch.bfloat16), t, text.to(dev), return_dict=False, parallel=par_sp)[0]
    res["sp_layout"] = [par_sp.cfg_size, par_sp.sp_size]
    res["sp_max_abs_diff"] = (single.float() - sharded.float()).abs().max().item()

    # 1b. the same with the exchange fused into the kernels over NVLink peer memory
    try:
        par_p2p = ParallelContext.create(use_cfg=False, use_p2p=True)
        fused = model(lat.to(dev, torch.bfloat16), t, text.to(dev), return_dict=False, parallel=par_p2p)[0]
        fused2 = model(lat.to(dev, torch.bfloat16), t, text.to(dev), return_dict=False, parallel=par_p2p)[0]
        res["p2p_max_abs_diff"] = (single.float() - fused.float()).abs().max().item()
        res["p2p_repeat_max_abs_diff"] = (fused2.float() - fused.float()).abs().max().item()
    except Exception as e:  # noqa
        import traceback

        res["p2p_error"] = repr(e)[:300] + " | " + traceback.format_exc()[-600:]

    # 2. CFG x SP denoise loop
    par = ParallelContext.create(use_cfg=True)

    def run(p):
        sch = UniPCMultistepScheduler(shift=3.0)
        sch.set_timesteps(4, device=dev)
        tr = denoise.DenoiseTrace()
        out = denoise.moe_denoise(timesteps=sch.timesteps, latents=lat.to(dev), scheduler=sch,
                                  high_noise_transformer=model, low_noise_transformer=model, boundary_timestep=875.0,
                                  guidance_scale=[4.0, 3.0], transformer_kwargs=dict(encoder_hidden_states=text.to(dev)),
                                  unconditional_transformer_kwargs=dict(encoder_hidden_states=neg.to(dev)), parallel=p,
                                  trace=tr)
        return out, tr

    a, tra = run(par)
    b, trb = run(ParallelContext.single())
    res["cfg_layout"] = [par.cfg_size, par.sp_size]
    res["denoise_max_abs_diff"] = (a - b).abs().max().item()
    res["trace_equal"] = tra.timesteps == trb.timesteps and tra.expert == trb.expert

    # 3. tile-parallel VAE decode
    wv = wan_vae.make_weights(base_dim=32, seed=7)
    vae = AutoencoderKLWan(WanVAEConfig(base_dim=32))
    vae.load_state_dict(wv, device=dev)
    vae.enable_tiling()
    z = torch.randn(1, 16, 3, 56, 40, generator=torch.Generator().manual_seed(5)).to(dev, torch.bfloat16)
    par_all = ParallelContext.create(use_cfg=False)
    v1 = vae.decode(z, return_dict=False)[0]
    v2 = vae.decode(z, return_dict=False, parallel=par_all)[0]
    res["vae_tiles"] = len(vae.tile_grid(56, 40))
    res["vae_max_abs_diff"] = (v1.float() - v2.float()).abs().max().item()

    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        print(json.dumps({"per_rank": gathered}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
