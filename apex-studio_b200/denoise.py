"""Wan denoise loops on the B200 backend.

Mirrors ``WanShared.moe_denoise`` / ``base_denoise`` (apps/api/src/engine/wan/shared/__init__.py:478-608,
:610-755) with the same keyword names: per step pick the expert by ``t >= boundary_timestep`` (:335-337,
both experts stay resident in the 180 GB of HBM instead of the reference's offload swap :341-462), pick the
guidance scale the same way (:464-476), run the conditional and the unconditional forward on
``latents.to(transformer_dtype)`` (:521), combine ``u + g * (c - u)`` in bf16 (:565), and advance the fp32
latents with ``scheduler.step`` (:569).

Multi-GPU (``parallel.ParallelContext``): the two CFG branches are independent given the same latents, so
with a CFG group of 2 each rank runs ONE branch and the branches are exchanged with a single all-gather of
the [B,16,F,H,W] bf16 prediction per step; every rank then takes the identical (deterministic, fp32)
scheduler step, so latents never need a broadcast.  Inside a forward the token axis may additionally be
sharded over a sequence-parallel group (wan/model.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence, Union

import torch

from . import ops
from .parallel import ParallelContext


@dataclass
class DenoiseTrace:
    """Integer decisions of a run -- compared with ``==`` against the oracle / reference in the tests."""
    timesteps: List[int] = field(default_factory=list)
    expert: List[str] = field(default_factory=list)          # "high" | "low" per step
    guidance: List[float] = field(default_factory=list)


def select_expert_is_high(t: Union[torch.Tensor, int, float], boundary_timestep) -> bool:
    """engine/wan/shared/__init__.py:335-337."""
    return boundary_timestep is not None and bool(t >= boundary_timestep)


def select_guidance_scale(t, boundary_timestep, guidance_scale: Union[float, Sequence[float]]) -> float:
    """engine/wan/shared/__init__.py:464-476."""
    if isinstance(guidance_scale, (list, tuple)):
        return float(guidance_scale[0] if select_expert_is_high(t, boundary_timestep) else guidance_scale[1])
    return float(guidance_scale)


class PreviewRenderer:
    """Per-step preview decode (reference: ``render_on_step`` in the denoise loops,
    engine/wan/shared/__init__.py:580-586 -> ``BaseEngine._render_step`` base_engine.py:2927-2943, which decodes the current
    latents with the VAE on the denoise stream and hands PIL frames to the callback, stalling the loop for the whole decode).

    Here the decode + frame hand-off (``vae.decode`` -> ``frames_to_uint8`` -> pinned host copy) is enqueued on a SIDE stream
    that only waits for the latents of the finished step, so the next step's kernels are never blocked by host work, and the
    callback is delivered with a uint8 ``[T, H, W, 3]`` numpy array once the copy has landed (at a later render point or at
    ``finish()``).  ``decode_fn(latents) -> uint8 tensor [T,H,W,3]`` is injected so that any VAE of this package fits."""

    def __init__(self, decode_fn: Callable[[torch.Tensor], torch.Tensor], callback: Callable, interval: int = 3):
        if interval < 1:
            raise ValueError("render_on_step_interval must be >= 1")
        self.decode_fn, self.callback, self.interval = decode_fn, callback, int(interval)
        self._side: Optional[torch.cuda.Stream] = None
        self._pending: List = []
        self.rendered_steps: List[int] = []

    def wants(self, i: int, total: int) -> bool:
        """The reference's condition (engine/wan/shared/__init__.py:581-584)."""
        return ((i + 1) % self.interval == 0 or i == 0) and i != total - 1

    def _deliver(self, block: bool) -> None:
        keep = []
        for ev, host, step in self._pending:
            if ev is None or block or ev.query():
                if ev is not None:
                    ev.synchronize()
                self.callback(host.numpy())
            else:
                keep.append((ev, host, step))
        self._pending = keep

    def maybe_render(self, i: int, total: int, latents: torch.Tensor) -> bool:
        self._deliver(block=False)
        if not self.wants(i, total):
            return False
        self.rendered_steps.append(i)
        if not latents.is_cuda:                                   # host-logic tests: run inline
            self._pending.append((None, self.decode_fn(latents).cpu(), i))
            self._deliver(block=True)
            return True
        if self._side is None:
            self._side = torch.cuda.Stream()
        snap = latents.detach().clone()
        self._side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._side):
            frames = self.decode_fn(snap)
            host = torch.empty(frames.shape, dtype=frames.dtype, pin_memory=True)
            host.copy_(frames, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._side)
        snap.record_stream(self._side)
        self._pending.append((ev, host, i))
        return True

    def finish(self) -> None:
        self._deliver(block=True)


def wan_preview_decode_fn(vae) -> Callable[[torch.Tensor], torch.Tensor]:
    """latents fp32 [1,16,F,H,W] -> uint8 [T,H,W,3]: ``BaseEngine.vae_decode`` (base_engine.py:2030-2059: denormalise, cast to the
    VAE dtype, forced tiling) followed by the frame hand-off kernel."""
    from .vae.wan import frames_to_uint8

    def fn(latents: torch.Tensor) -> torch.Tensor:
        z = vae.denormalize_latents(latents).to(torch.bfloat16)
        vae.enable_tiling()
        return frames_to_uint8(vae.decode(z, return_dict=False)[0][0].contiguous())

    return fn


@torch.inference_mode()
def moe_denoise(*, timesteps: torch.Tensor, latents: torch.Tensor, scheduler, high_noise_transformer,
                low_noise_transformer=None, boundary_timestep=None, guidance_scale: Union[float, Sequence[float]] = 5.0,
                transformer_kwargs: Optional[Dict[str, Any]] = None,
                unconditional_transformer_kwargs: Optional[Dict[str, Any]] = None, use_cfg_guidance: bool = True,
                transformer_dtype=torch.bfloat16, extra_step_kwargs: Optional[Dict[str, Any]] = None,
                denoise_progress_callback: Optional[Callable] = None, parallel: Optional[ParallelContext] = None,
                trace: Optional[DenoiseTrace] = None, render_on_step: bool = False,
                render_on_step_callback: Optional[Callable] = None, render_on_step_interval: int = 3,
                preview_decode_fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None) -> torch.Tensor:
    """Returns the final fp32 latents.  ``low_noise_transformer=None`` gives ``base_denoise`` (single expert).
    ``render_on_step`` / ``render_on_step_callback`` / ``render_on_step_interval`` keep the reference's names (:496-503);
    ``preview_decode_fn`` (e.g. ``wan_preview_decode_fn(vae)``) supplies the decode that the engine's ``_render_step`` does."""
    transformer_kwargs = dict(transformer_kwargs or {})
    uncond_kwargs = dict(unconditional_transformer_kwargs or {})
    transformer_kwargs.pop("encoder_hidden_states_image", None)
    uncond_kwargs.pop("encoder_hidden_states_image", None)
    extra_step_kwargs = extra_step_kwargs or {}
    do_cfg = bool(use_cfg_guidance and uncond_kwargs)
    par = parallel or ParallelContext.single()
    total = len(timesteps)
    preview = None
    if render_on_step and render_on_step_callback is not None:
        if preview_decode_fn is None:
            raise ValueError("render_on_step needs `preview_decode_fn` (e.g. denoise.wan_preview_decode_fn(vae))")
        preview = PreviewRenderer(preview_decode_fn, render_on_step_callback, render_on_step_interval)
    # host copy of the schedule, ONE device read per run: the expert switch `t >= boundary` and the guidance choice are host
    # decisions (the reference's `t >= boundary` on a device tensor synchronises every step, :335-337), which also keeps
    # the per-step path free of host<->device syncs (CUDA-graph friendly)
    ts_host = timesteps.tolist() if isinstance(timesteps, torch.Tensor) else [float(x) for x in timesteps]
    for i, t in enumerate(timesteps):
        latent_model_input = latents.to(transformer_dtype)
        timestep = t.expand(latents.shape[0])
        t_host = ts_host[i]
        high = select_expert_is_high(t_host, boundary_timestep)
        transformer = high_noise_transformer if (high or low_noise_transformer is None) else low_noise_transformer
        g = select_guidance_scale(t_host, boundary_timestep, guidance_scale)
        if trace is not None:
            trace.timesteps.append(int(t_host))
            trace.expert.append("high" if (high or low_noise_transformer is None) else "low")
            trace.guidance.append(g)

        if do_cfg and par.cfg_size == 2:
            # one branch per CFG rank, one all-gather of the predictions
            kw = transformer_kwargs if par.cfg_rank == 0 else uncond_kwargs
            mine = transformer(hidden_states=latent_model_input, timestep=timestep, return_dict=False,
                               parallel=par, **kw)[0]
            cond, uncond = par.exchange_cfg(mine)
            noise_pred = ops.cfg_combine(cond, uncond, g)
        else:
            cond = transformer(hidden_states=latent_model_input, timestep=timestep, return_dict=False,
                               parallel=par, **transformer_kwargs)[0]
            if do_cfg:
                uncond = transformer(hidden_states=latent_model_input, timestep=timestep, return_dict=False,
                                     parallel=par, **uncond_kwargs)[0]
                noise_pred = ops.cfg_combine(cond, uncond, g)
            else:
                noise_pred = cond
        latents = scheduler.step(noise_pred, t, latents, **extra_step_kwargs, return_dict=False)[0]
        if preview is not None and par.rank == 0:
            preview.maybe_render(i, total, latents)
        if denoise_progress_callback is not None:
            denoise_progress_callback(float(i + 1) / float(total), f"Denoising step {i + 1}/{total}")
    if preview is not None:
        preview.finish()
    return latents


def base_denoise(**kwargs) -> torch.Tensor:
    """Single-expert loop (engine/wan/shared/__init__.py:610-755): ``transformer=`` instead of the expert pair."""
    transformer = kwargs.pop("transformer")
    kwargs.pop("low_noise_transformer", None)
    kwargs["boundary_timestep"] = None
    return moe_denoise(high_noise_transformer=transformer, low_noise_transformer=None, **kwargs)


@torch.inference_mode()
def flux_denoise(*, latents: torch.Tensor, timesteps: torch.Tensor, scheduler, transformer, prompt_embeds: torch.Tensor,
                 pooled_prompt_embeds: torch.Tensor, latent_ids: torch.Tensor, text_ids: torch.Tensor,
                 guidance: Optional[torch.Tensor] = None, negative_prompt_embeds: Optional[torch.Tensor] = None,
                 negative_pooled_prompt_embeds: Optional[torch.Tensor] = None, negative_text_ids: Optional[torch.Tensor] = None,
                 true_cfg_scale: float = 1.0, use_cfg_guidance: bool = False,
                 denoise_progress_callback: Optional[Callable] = None) -> torch.Tensor:
    """``FluxShared.base_denoise`` (apps/api/src/engine/flux/shared.py:504-620) with the same keyword names: per step one
    forward on the packed latents [B, S_img, 64] at ``timestep / 1000`` (plus the negative-prompt forward and
    ``neg + s * (pos - neg)`` when true CFG is on, :561-582), then ``scheduler.step``.  FLUX.1-dev is guidance-distilled:
    ``guidance`` is embedded (:150-156 of t2i.py) and true CFG is normally off."""
    total = len(timesteps)
    for i, t in enumerate(timesteps):
        timestep = t.expand(latents.shape[0]).to(latents.dtype)
        kw = dict(hidden_states=latents, timestep=timestep / 1000, guidance=guidance, img_ids=latent_ids, return_dict=False)
        noise_pred = transformer(pooled_projections=pooled_prompt_embeds, encoder_hidden_states=prompt_embeds,
                                 txt_ids=text_ids, **kw)[0]
        if use_cfg_guidance:
            neg = transformer(pooled_projections=negative_pooled_prompt_embeds, encoder_hidden_states=negative_prompt_embeds,
                              txt_ids=negative_text_ids, **kw)[0]
            noise_pred = ops.cfg_combine(noise_pred, neg, float(true_cfg_scale))
        latents = scheduler.step(noise_pred, t, latents, return_dict=False)[0]
        if denoise_progress_callback is not None:
            denoise_progress_callback(float(i + 1) / float(total), f"Denoise {i + 1}/{total}")
    return latents


@torch.inference_mode()
def hy15_denoise(*, timesteps: torch.Tensor, latents: torch.Tensor, scheduler, transformer, cond_latents_concat: torch.Tensor,
                 mask_concat: torch.Tensor, image_embeds: torch.Tensor, cond_kwargs: Dict[str, Any],
                 uncond_kwargs: Optional[Dict[str, Any]] = None, guidance_scale: float = 6.0,
                 do_classifier_free_guidance: bool = True, guidance_rescale: float = 0.0,
                 denoise_progress_callback: Optional[Callable] = None, parallel: Optional[ParallelContext] = None) -> torch.Tensor:
    """The HunyuanVideo-1.5 denoise loop (apps/api/src/engine/hunyuanvideo15/t2v.py:234-338) with the same names: per step
    ``latent_model_input = cat([latents, cond_latents_concat, mask_concat], dim=1)`` (:239-241), the timestep in the latent
    dtype (:243-245), the unconditional and the conditional forward (:262-289), ``uncond + g * (text - uncond)`` (:291-293)
    and ``scheduler.step`` (:320-322).  ``cond_kwargs`` / ``uncond_kwargs`` carry ``encoder_hidden_states``,
    ``encoder_attention_mask``, ``encoder_hidden_states_2``, ``encoder_attention_mask_2`` (:248-261).

    Multi-GPU: with a CFG group of 2 each rank runs ONE branch and the [B,32,F,H,W] bf16 predictions are exchanged with a
    single all-gather per step; inside a forward the latent tokens may be sharded over the sequence-parallel group."""
    if guidance_rescale and guidance_rescale > 0.0:
        raise ValueError("guidance_rescale > 0 is not implemented on the b200 path")
    par = parallel or ParallelContext.single()
    do_cfg = bool(do_classifier_free_guidance)
    if do_cfg and not uncond_kwargs:
        raise ValueError("CFG requested (guidance_scale > 1.0) but no negative prompt / negative prompt embeds were provided.")
    total = len(timesteps)
    for i, t in enumerate(timesteps):
        latent_model_input = torch.cat([latents, cond_latents_concat, mask_concat], dim=1)
        timestep = t.expand(latent_model_input.shape[0]).to(latent_model_input.dtype)
        common = dict(hidden_states=latent_model_input, image_embeds=image_embeds, timestep=timestep, return_dict=False,
                      parallel=par)
        if do_cfg and par.cfg_size == 2:
            mine = transformer(**common, **(cond_kwargs if par.cfg_rank == 0 else uncond_kwargs))[0]
            cond, uncond = par.exchange_cfg(mine)
            noise_pred = ops.cfg_combine(cond, uncond, float(guidance_scale))
        elif do_cfg:
            uncond = transformer(**common, **uncond_kwargs)[0]
            cond = transformer(**common, **cond_kwargs)[0]
            noise_pred = ops.cfg_combine(cond, uncond, float(guidance_scale))
        else:
            noise_pred = transformer(**common, **cond_kwargs)[0]
        latents = scheduler.step(noise_pred, t, latents, return_dict=False)[0]
        if denoise_progress_callback is not None:
            denoise_progress_callback(float(i + 1) / float(max(total, 1)), f"Denoising step {i + 1}/{total}")
    return latents
