#!/usr/bin/env python
"""Secondary benchmark (BASELINE.json configs[4]): HunyuanVideo-1.5 720p x 129 frames on B200s: latent grid 33 x 45 x 80 =
118,800 tokens + 1985 condition tokens (1000 MLLM + 256 ByT5 + 729 image), d = 2048, 16 heads, 54 dual-stream blocks (91 %
attention).  One denoise step = unconditional + conditional forward + CFG combine + flow-match Euler update
(engine/hunyuanvideo15/t2v.py:234-338).

    python scripts/bench_hy15.py [--steps K] [--warmup W] [--layers 54]                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_hy15.py
N > 1: CFG pair x latent-token shards (Ulysses over NCCL; the text stream is replicated) -- strong scaling of ONE job.
Prints one JSON line (rank 0): steps/s (CUDA events, max over ranks), algorithmic TFLOP/s against the measured bf16 peak."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--layers", type=int, default=54)
    ap.add_argument("--p2p", action="store_true", help="fused NVLink peer-memory exchange (parallel.JointPeerExchange) instead of NCCL")
    a = ap.parse_args()
    from apex_studio_b200 import denoise, ops
    from apex_studio_b200.hunyuanvideo15 import HunyuanVideo15Config, HunyuanVideo15Transformer3DModel
    from apex_studio_b200.parallel import ParallelContext
    from apex_studio_b200.scheduler import FlowMatchEulerDiscreteScheduler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    par = ParallelContext.create(use_cfg=True, use_p2p=a.p2p) if world > 1 else ParallelContext.single()
    m = HunyuanVideo15Transformer3DModel(HunyuanVideo15Config(num_layers=a.layers)).init_random_weights(dev)
    g = torch.Generator(device=dev).manual_seed(42)
    bf = torch.bfloat16
    lat = torch.randn(1, 32, 33, 45, 80, generator=g, device=dev).to(bf)
    cond_lat = torch.zeros(1, 32, 33, 45, 80, device=dev, dtype=bf)
    mask = torch.zeros(1, 1, 33, 45, 80, device=dev, dtype=bf)
    text = torch.randn(1, 1000, 3584, generator=g, device=dev).to(bf)
    neg = torch.randn(1, 1000, 3584, generator=g, device=dev).to(bf)
    text2 = torch.randn(1, 256, 1472, generator=g, device=dev).to(bf)
    img = torch.zeros(1, 729, 1152, device=dev, dtype=bf)      # t2v: all-zero image embeds (t2v.py:196-202)
    m1, m2 = torch.ones(1, 1000), torch.ones(1, 256)
    m1[:, 300:], m2[:, 64:] = 0, 0
    kw = dict(encoder_hidden_states=text, encoder_attention_mask=m1, encoder_hidden_states_2=text2, encoder_attention_mask_2=m2)
    n_sched = 30
    sch = FlowMatchEulerDiscreteScheduler(use_dynamic_shifting=False, shift=7.0)
    ts = sch.set_timesteps(n_sched, device=dev, sigmas=np.linspace(1.0, 0.0, n_sched + 1)[:-1])

    def steps(lo, n):
        sch._step_index = lo
        return denoise.hy15_denoise(timesteps=ts[lo:lo + n], latents=lat, scheduler=sch, transformer=m, cond_latents_concat=cond_lat,
                                    mask_concat=mask, image_embeds=img, cond_kwargs=kw, uncond_kwargs=dict(kw, encoder_hidden_states=neg),
                                    guidance_scale=6.0, parallel=par)

    y = steps(0, a.warmup)
    ops.launch_count = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    from bench import ClockSampler   # nvidia-smi clocks / throttle reasons sampled during the timed region (rank 0's GPU)
    clk = ClockSampler(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clk.start()
    e0.record()
    y = steps(a.warmup, a.steps)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = clk.stop()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    S, d = 118800 + 1985, 2048
    fl = 2 * a.layers * (24.0 * S * d * d + 4.0 * S * S * d)
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk)).get("bf16_tflops_sustained", 1376.1) if os.path.exists(pk) else 1376.1
    if rank == 0:
        print(json.dumps({"metric": "denoise_steps_per_sec", "workload": "HunyuanVideo-1.5 720p x 129f (118800 latent + 1985 condition "
                          "tokens, d=2048, 16 heads, %d dual-stream blocks), CFG on: 2 forwards + combine + Euler per step" % a.layers,
                          "value": 1000.0 / ms, "unit": "steps/s", "ms_per_step": ms, "n_gpus": world,
                          "parallelism": f"cfg{par.cfg_size} x sp{par.sp_size}", "exchange": "fused peer memory" if (a.p2p and par.sp_size > 1) else "nccl", "steps": a.steps, "warmup": a.warmup, "dtype": "bf16",
                          "data": "synthetic", "clocks": clocks, "algorithmic_flops_per_step": fl, "tflops": fl / ms / 1e9,
                          "tflops_per_gpu": fl / ms / 1e9 / world, "frac_of_peak_per_gpu": fl / ms / 1e9 / world / peak, "peak": peak,
                          "gpu_launches_per_step_rank0": ops.launch_count // a.steps, "finite": bool(torch.isfinite(y).all()),
                          "parameter_gb": m.parameter_bytes() / 1e9}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
