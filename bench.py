#!/usr/bin/env python
"""Headline benchmark: denoise-steps/sec of Wan-2.2 A14B t2v, 720p x 81 frames, CFG on (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = what `moe_denoise` does per timestep (apps/api/src/engine/wan/shared/__init__.py:514-569): the
conditional and the unconditional DiT forward over latents [1,16,21,90,160] (S = 75,600 tokens, 40 layers,
d = 5120, 40 heads x 128, ffn 13,824, 512 text tokens), the CFG combine and the UniPC scheduler step.
Synthetic latents / text embeddings / random-init bf16 weights of the A14B architecture (both experts
resident); the first W+K(+e2e) timesteps of a 50-step schedule are run (with shift 3 only the first ~15 timesteps are
>= 875, so a longer run crosses the high-noise -> low-noise expert switch; both experts are resident either way).

Printed JSON (one line, rank 0): `value` = steps/sec with inputs resident in HBM (CUDA events, max over ranks);
`e2e` = the same step driven from HOST buffers through the public API (pinned-host -> device copy of latents
and text embeddings, device -> host read of the new latents inside the timed region); `roofline` = the
self-attention kernel (dominant: 72 % of the FLOPs) timed per launch with CUDA events on the launching
stream, against the measured sustained bf16 peak in MEASURED_PEAKS.json; `cpu_baseline` = the oracle port of
the reference's PyTorch path on this box's host cores over a bounded sample, extrapolated (labelled).
N > 1: CFG pair x token/head shards (apex-studio_b200/parallel.py); strong scaling of ONE job.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "denoise_steps_per_sec"
UNIT = "steps/s"
LATENT_SHAPE = (1, 16, 21, 90, 160)
TEXT_LEN, TEXT_DIM = 512, 4096
S_TOKENS, DIM, HEADS, FFN, LAYERS = 75600, 5120, 40, 13824, 40
# algorithmic FLOPs (SURVEY.md section 8d / BASELINE.md section 3)
FLOPS_ATTN_LAUNCH = 4.0 * S_TOKENS * S_TOKENS * DIM                 # 1.17051e14 per self-attention launch
FLOPS_FORWARD = LAYERS * (8.0 * S_TOKENS * DIM * DIM + FLOPS_ATTN_LAUNCH
                          + (4.0 * S_TOKENS * DIM * DIM + 4.0 * TEXT_LEN * DIM * DIM + 4.0 * S_TOKENS * TEXT_LEN * DIM)
                          + 4.0 * S_TOKENS * DIM * FFN)                # 6.5233e15
FLOPS_STEP = 2.0 * FLOPS_FORWARD                                     # 1.30466e16 (cond + uncond)
WORKLOAD = "Wan-2.2 A14B t2v 720p x 81f (latent 1x16x21x90x160, S=75600), CFG on, UniPC 50-step schedule"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": d.get("bf16_tflops_sustained", 1376.1), "hbm_gbs": d.get("hbm_gbs", 6552.6), "src": "measured"}
    return {"tflops": 1400.0, "hbm_gbs": 6650.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu_index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------
# CPU baseline: oracle port of the reference's PyTorch path, bounded sample, host cores
# ----------------------------------------------------------------------------------------------------------
def cpu_reference_sample(n_tok: int = 256, repeats: int = 1):
    """Times ONE WanTransformerBlock of the A14B dimensions (oracle/wan_dit.py, bf16 like the reference's torch
    path) on `n_tok` query tokens: (a) every token-wise op of the block on n_tok tokens (LN/modulate, q/k/v and
    out projections, cross-attention to 512 text tokens, FFN, gates), (b) the self-attention core for n_tok
    queries against all 75,600 keys x 40 heads.  Both scale linearly with the number of query tokens, so
    step time = (t_a + t_b) * (75600 / n_tok) * 40 layers * 2 forwards  (EXTRAPOLATED, labelled in the output)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import wan_dit

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    bf = torch.bfloat16
    w = wan_dit.make_weights(dim=DIM, heads=HEADS, ffn_dim=FFN, num_layers=1, text_dim=64, seed=1234, dtype=bf)
    g = torch.Generator().manual_seed(0)
    h = torch.randn(1, n_tok, DIM, generator=g).to(bf)
    ctx = torch.randn(1, TEXT_LEN, DIM, generator=g).to(bf)
    temb6 = torch.randn(1, 6, DIM, generator=g).to(bf)
    freqs = wan_dit.rope_table(128, (1, 16, n_tok // 16))
    q = torch.randn(1, HEADS, n_tok, 128, generator=g).to(bf)
    k = torch.randn(1, HEADS, S_TOKENS, 128, generator=g).to(bf)
    v = torch.randn(1, HEADS, S_TOKENS, 128, generator=g).to(bf)
    with torch.inference_mode():
        wan_dit.block_forward(h[:, :16], ctx, temb6, freqs[:16], w, "blocks.0", HEADS)  # warm-up
        tas, tbs = [], []
        for _ in range(repeats):
            t0 = time.perf_counter()
            wan_dit.block_forward(h, ctx, temb6, freqs, w, "blocks.0", HEADS)
            t1 = time.perf_counter()
            wan_dit.sdpa(q, k, v)
            t2 = time.perf_counter()
            tas.append(t1 - t0)
            tbs.append(t2 - t1)
    ta, tb = statistics.median(tas), statistics.median(tbs)     # median of the repeats: the host cores are shared
    scale = S_TOKENS / n_tok
    step_s = (ta + tb) * scale * LAYERS * 2
    sample = (f"1 WanTransformerBlock (d=5120, ffn=13824, 40 heads, bf16, torch CPU) on {n_tok} query tokens: token-wise ops "
              f"{ta:.2f}s + self-attention core vs all 75600 keys {tb:.2f}s; EXTRAPOLATED x{scale:.1f} tokens x40 layers "
              f"x2 forwards = {step_s:.0f}s per denoise step")
    return {"value": 1.0 / step_s, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
            "measured_s": {"tokenwise": ta, "attention_core": tb}, "n_tok": n_tok}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    vals = []
    for _ in range(max(1, args.warmup // 3)):
        cpu_reference_sample(n_tok=64)
    for _ in range(args.steps):
        vals.append(cpu_reference_sample(n_tok=4096, repeats=3))   # ~5-8 s of host work per step, median of 3
    v = statistics.mean(x["value"] for x in vals)
    base = dict(vals[-1], value=v)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "arm": "oracle port of the reference PyTorch path on host cores; each "
                       "step = bounded sample (see cpu_baseline.sample), extrapolated to the full step"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    emit(line)


# ----------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch.distributed as dist

    from apex_studio_b200 import denoise, ops
    from apex_studio_b200.parallel import ParallelContext
    from apex_studio_b200.scheduler import UniPCMultistepScheduler
    from apex_studio_b200.wan import WanConfig, WanTransformer3DModel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the single JSON line (no "NCCL version ..." banner)
        dist.init_process_group("nccl", device_id=dev)
    # Ulysses exchange fused into the kernels over NVLink peer memory (default; measured 6.7 % faster than the NCCL
    # all-to-all at N=4 and bit-identical); B200_P2P=0 selects the NCCL exchange.
    use_p2p = os.environ.get("B200_P2P", "1") == "1"
    par = ParallelContext.create(use_cfg=True, use_p2p=use_p2p) if world > 1 else ParallelContext.single()

    cfg = WanConfig(num_layers=args.layers)
    high = WanTransformer3DModel(cfg).init_random_weights(dev, seed=1234)
    low = WanTransformer3DModel(cfg).init_random_weights(dev, seed=4321)
    if args.cuda_graph and world == 1:
        high.enable_cuda_graph()
        low.enable_cuda_graph()
    sch = UniPCMultistepScheduler(shift=3.0)
    sch.set_timesteps(50, device=dev)
    boundary = 0.875 * sch.num_train_timesteps

    # host-side (pinned) inputs, exactly what the engine holds before a step
    g = torch.Generator().manual_seed(42)
    lat_host = torch.randn(LATENT_SHAPE, generator=g, dtype=torch.float32).pin_memory()
    pos_host = torch.randn(1, TEXT_LEN, TEXT_DIM, generator=torch.Generator().manual_seed(43)).bfloat16().pin_memory()
    neg_host = torch.randn(1, TEXT_LEN, TEXT_DIM, generator=torch.Generator().manual_seed(44)).bfloat16().pin_memory()
    out_host = torch.empty(LATENT_SHAPE, dtype=torch.float32).pin_memory()
    h2d_bytes = lat_host.numel() * 4 + pos_host.numel() * 2 + neg_host.numel() * 2
    d2h_bytes = out_host.numel() * 4

    # per-launch timing of the dominant kernel (self-attention) with events on the launching stream
    attn_events = []
    orig_attention = ops.attention

    def timed_attention(q, k, v, softmax_scale=None, out=None, **kw):
        if q.shape[2] == k.shape[2] and q.shape[2] >= 4096 and timed_attention.on:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = orig_attention(q, k, v, softmax_scale, out, **kw)
            e.record()
            attn_events.append((s, e, q.shape[1], q.shape[2]))
            return r
        return orig_attention(q, k, v, softmax_scale, out, **kw)

    timed_attention.on = False
    ops.attention = timed_attention

    orig_scatter = ops.attention_scatter   # sequence-parallel form of the same kernel (fused heads->tokens exchange)

    def timed_scatter(q, k, v, *a, **kw):
        if q.shape[2] == k.shape[2] and q.shape[2] >= 4096 and timed_attention.on:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            orig_scatter(q, k, v, *a, **kw)
            e.record()
            attn_events.append((s, e, q.shape[1], q.shape[2]))
            return None
        return orig_scatter(q, k, v, *a, **kw)

    ops.attention_scatter = timed_scatter

    state = {"latents": lat_host.to(dev), "pos": pos_host.to(dev), "neg": neg_host.to(dev), "i": 0,
             "host_in": lat_host, "host_out": out_host}

    def one_step(from_host: bool):
        i = state["i"]
        if from_host:  # this step's inputs come from pinned host memory
            state["latents"] = state["host_in"].to(dev, non_blocking=True)
            state["pos"].copy_(pos_host, non_blocking=True)
            state["neg"].copy_(neg_host, non_blocking=True)
        t = sch.timesteps[i:i + 1]
        new = denoise.moe_denoise(timesteps=t, latents=state["latents"], scheduler=sch, high_noise_transformer=high,
                                  low_noise_transformer=low, boundary_timestep=boundary, guidance_scale=[4.0, 3.0],
                                  transformer_kwargs=dict(encoder_hidden_states=state["pos"]),
                                  unconditional_transformer_kwargs=dict(encoder_hidden_states=state["neg"]),
                                  parallel=par)
        state["i"] = i + 1
        if from_host:  # ... and its result is read back to the host, which owns the latents between steps
            state["host_out"].copy_(new, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            state["host_in"], state["host_out"] = state["host_out"], state["host_in"]
        else:
            state["latents"] = new

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        one_step(False)

    # ---- device-resident timed region: EXACTLY K steps
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = ops.launch_count
    timed_attention.on = True
    torch.cuda.profiler.start()   # `ncu --profile-from-start off` captures exactly the timed region
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        one_step(False)
    ev1.record()
    barrier()
    torch.cuda.profiler.stop()
    timed_attention.on = False
    launches = ops.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_step = ms_total / args.steps

    # result checksum: the fp32 latents after the W + K device-resident steps, byte for byte.  Every N runs the same seeded
    # job, so the N = 1 / 2 / 4 / 8 lines of a scaling run can be compared for the claimed bit-identity of the sharded paths.
    latents_sha = hashlib.sha256(state["latents"].detach().float().cpu().contiguous().numpy().tobytes()).hexdigest()

    attn_ms = [s.elapsed_time(e) for s, e, _, _ in attn_events]
    heads_per_launch = attn_events[0][2] if attn_events else HEADS
    attn_avg = statistics.mean(attn_ms) if attn_ms else float("nan")

    # ---- end-to-end region: host buffers, copies inside the timed region
    e2e_steps = args.e2e_steps if args.e2e_steps is not None else args.steps
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        one_step(True)
    e1.record()
    barrier()
    e2e_ms_step = max_over_ranks(e0.elapsed_time(e1)) / max(e2e_steps, 1)

    # ---- VAE decode of the same latent (second half of the frames/sec metric); tiles dealt over all ranks
    vae_info = None
    if not args.no_vae:
        try:
            from apex_studio_b200.vae import AutoencoderKLWan

            del high, low, state
            torch.cuda.empty_cache()
            vae = AutoencoderKLWan().init_random_weights(dev, seed=7)
            vae.enable_tiling()
            zlat = vae.denormalize_latents(torch.randn(LATENT_SHAPE, generator=torch.Generator().manual_seed(42))
                                           .to(dev)).to(torch.bfloat16)
            par_v = ParallelContext.create(use_cfg=False) if world > 1 else None
            l0 = ops.launch_count
            vae.decode(zlat, parallel=par_v)                      # warm-up (allocator, TMA descriptors, clocks)
            vae_launches = ops.launch_count - l0
            barrier()
            v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            v0.record()
            frames = vae.decode(zlat, parallel=par_v)[0]
            v1.record()
            barrier()
            vae_ms = max_over_ranks(v0.elapsed_time(v1))
            vae_info = {"ms": vae_ms, "frames": int(frames.shape[2]), "frames_per_sec": frames.shape[2] / (vae_ms * 1e-3),
                        "tiles": len(vae.tile_grid(LATENT_SHAPE[3], LATENT_SHAPE[4])), "launches": vae_launches,
                        "algorithmic_tflops_untiled": 6.39e14 / (vae_ms * 1e-3) / 1e12,
                        "finite": bool(torch.isfinite(frames.float()).all().item()),
                        "frames_sha256": hashlib.sha256(frames.detach().view(torch.int16).cpu().contiguous().numpy().tobytes()).hexdigest()}
        except Exception as e:  # the DiT line must survive a VAE problem
            vae_info = {"error": repr(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    attn_flops = FLOPS_ATTN_LAUNCH * heads_per_launch / HEADS
    achieved = attn_flops / (attn_avg * 1e-3) / 1e12 if attn_ms else None
    traffic = None
    prof = os.path.join(ROOT, "profiles", "attn_ncu_summary.json")
    if os.path.exists(prof):
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from the committed `ncu --set full` capture
            d = json.load(open(prof))
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
            traffic = sum(float(d[k]["value"]) * mult[d[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        except Exception:
            traffic = None
    full_model = args.layers == LAYERS
    step_flops = FLOPS_STEP * args.layers / LAYERS
    line = {
        "metric": METRIC, "value": 1000.0 / ms_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD if full_model else f"REDUCED ({args.layers} layers, debug only) " + WORKLOAD,
                   "layers": args.layers, "experts_resident": 2, "cuda_graph": bool(args.cuda_graph and world == 1), "guidance_scale": [4.0, 3.0], "flow_shift": 3.0,
                   "parallelism": f"cfg{par.cfg_size} x sp{par.sp_size}" + (" (peer-memory fused exchange)" if (
                       par.use_p2p and par.sp_size > 1) else (" (NCCL all-to-all)" if par.sp_size > 1 else "")), "l2": "inputs larger than L2 (0.77 GB activations)",
                   "timesteps_run": f"first {args.warmup + args.steps + e2e_steps} of 50"},
        "roofline": {"bound": "tensor", "kernel": "attn_fwd_kernel (self-attention, S=75600)", "achieved": achieved,
                     "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": (achieved / peaks["tflops"]) if achieved else None,
                     "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write, profiles/attn_ncu_summary.json; "
                     "algorithmic q+k+v+o = 3.10e9)", "peak_source": peaks["src"] + " sustained bf16 (MEASURED_PEAKS.json)",
                     "launches_timed": len(attn_ms), "avg_launch_ms": attn_avg,
                     "kernel_share_of_step": (sum(attn_ms) / args.steps) / ms_step if attn_ms else None,
                     "algorithmic_flops_per_launch": attn_flops},
        "roofline_step": {"achieved": step_flops / (ms_step * 1e-3) / 1e12 * (1.0 if world == 1 else 1.0),
                          "unit": "TFLOP/s (whole job)", "per_gpu": step_flops / (ms_step * 1e-3) / 1e12 / world,
                          "frac_per_gpu": step_flops / (ms_step * 1e-3) / 1e12 / world / peaks["tflops"],
                          "algorithmic_flops_per_step": step_flops},
        "e2e": {"value": 1000.0 / e2e_ms_step, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps, "ms_per_step": e2e_ms_step},
        "gpu_launches": launches, "clocks": clocks,
        "latents_sha256": latents_sha, "latents_sha256_of": f"fp32 latents after {args.warmup + args.steps} steps (seeded; identical for every N)",
        "frames_per_sec_50step_denoise_only": 81.0 / (50 * ms_step * 1e-3),
        "vae_decode": vae_info,
        "frames_per_sec": (81.0 / (50 * ms_step * 1e-3 + vae_info["ms"] * 1e-3)) if (vae_info and "ms" in vae_info) else None,
    }
    if world == 1 and not args.no_reference_gpu:
        # "reference on 1 x B200" (BASELINE.md section 4): the reference's OWN WanTransformerBlock (unmodified files under
        # baseline/_ref; torch SDPA / flash-attn + cuBLAS + ATen) at the same shape, in a separate process, OUTSIDE every timed
        # region of this arm.  Reported, extrapolated per step (x 40 layers x 2 forwards), never mixed into `value`.
        try:
            torch.cuda.empty_cache()
            r = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "ref_gpu_block.py"), "--iters", "3"],
                               capture_output=True, text=True, timeout=600)
            rows = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            line["reference_gpu"] = json.loads(rows[-1]) if rows else {"error": (r.stderr or r.stdout)[-300:]}
            rg = line["reference_gpu"]
            if "steps_per_sec_extrapolated" in rg:
                rg["this_repo_over_reference_gpu"] = line["value"] / rg["steps_per_sec_extrapolated"]
        except Exception as e:
            line["reference_gpu"] = {"error": repr(e)[:300]}
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_reference_sample(n_tok=4096, repeats=3)
        except Exception as e:  # keep the GPU line even if the host is too small for the sample
            line["cpu_baseline"] = {"error": repr(e)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def emit(line: dict) -> None:
    """Write the ONE result line to the process's real stdout (see main(): fd 1 is pointed at stderr while the
    benchmark runs so that banners printed by native libraries, e.g. "NCCL version ...", cannot pollute it)."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, data)


def main():
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--layers", type=int, default=LAYERS, help="debug only: anything but 40 is labelled REDUCED")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-vae", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true")
    ap.add_argument("--cuda-graph", action="store_true", help="replay each expert's forward from one CUDA graph (N=1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
