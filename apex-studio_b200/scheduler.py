"""Flow-matching UniPC multistep scheduler (predictor-corrector, B(h) variants) for the Wan denoise loop.

Mirrors the scheduler the reference steps with (manifest: ``diffusers.UniPCMultistepScheduler`` with flow
sigmas, apps/api/manifest/video/wan-2.2-a14b-text-to-video-1.0.0.v1.yml:55-61; in-repo twin
apps/api/src/scheduler/unipc.py:19) -- same constructor names, ``set_timesteps`` / ``step`` /
``index_for_timestep`` / ``set_begin_index`` surface and the same state machine:

* integer quantities (bit-exact requirement): ``timesteps`` = trunc(sigma * 1000) as
  int64 (unipc.py:185-212), first-step index lookup with the "second match" rule (:624-636), ``_step_index``
  += 1 per step (:732), order warm-up ``min(order, len - idx, lower_order_nums + 1)`` (:709-719), corrector
  gate (:687-691);
* float quantities: the sigma schedule is built in float64 numpy and stored as float32, the B(h) coefficients
  are float32 scalars computed on the host in the reference's operation order, and the latent update is a
  handful of fp32 element-wise torch ops on the device (4.8 M elements per step at 720p x 81f -- the
  scheduler is O(numel) and stays in torch, as scoped in SURVEY.md).

Two sigma-schedule flavours:
  ``flavor="twin"``      apps/api/src/scheduler/unipc.py: linspace(sigma_max, sigma_min, n+1)[:-1] of the
                         *shifted* training schedule, shifted again in ``set_timesteps`` (:185-196);
  ``flavor="diffusers"`` upstream diffusers ``use_flow_sigmas=True``: flip(shift*s/(1+(shift-1)s))[:-1] with
                         s = 1 - linspace(1, 1/num_train_timesteps, n+1) (restated from the published
                         algorithm; diffusers is not vendored by the reference -- parity unpinned).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch


class SchedulerOutput(tuple):
    """What ``step(..., return_dict=True)`` returns: diffusers' ``SchedulerOutput`` exposes ``.prev_sample`` and, like every
    ``BaseOutput``, positional indexing (``out[0]``)."""

    def __new__(cls, prev_sample):
        return super().__new__(cls, (prev_sample,))

    @property
    def prev_sample(self):
        return self[0]


class UniPCMultistepScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, solver_order: int = 2, prediction_type: str = "flow_prediction",
                 shift: float = 1.0, predict_x0: bool = True, solver_type: str = "bh2", lower_order_final: bool = True,
                 disable_corrector: Sequence[int] = (), final_sigmas_type: str = "zero", flavor: str = "twin"):
        if prediction_type != "flow_prediction":
            raise ValueError(f"prediction_type given as {prediction_type} must be `flow_prediction` here")
        if solver_type not in ("bh1", "bh2"):
            if solver_type in ("midpoint", "heun", "logrho"):
                solver_type = "bh2"
            else:
                raise NotImplementedError(f"{solver_type} is not implemented for {self.__class__}")
        if flavor not in ("twin", "diffusers"):
            raise ValueError(f"unknown flavor {flavor}")
        if final_sigmas_type != "zero":
            raise ValueError("`final_sigmas_type` must be 'zero' for the flow schedule")
        self.num_train_timesteps = num_train_timesteps
        self.solver_order = solver_order
        self.shift = shift
        self.predict_x0 = predict_x0
        self.solver_type = solver_type
        self.lower_order_final = lower_order_final
        self.disable_corrector = list(disable_corrector)
        self.flavor = flavor
        self.config = self  # reference code reads scheduler.config.num_train_timesteps (t2v.py:177-181)

        # training schedule in float32 (unipc.py:112-123): sigma_k = 1 - alpha_k, shifted
        alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
        sig = torch.from_numpy(1.0 - alphas).to(dtype=torch.float32)
        sig = shift * sig / (1 + (shift - 1) * sig)
        self.sigmas = sig
        self.timesteps = sig * num_train_timesteps
        self.sigma_min = sig[-1].item()
        self.sigma_max = sig[0].item()
        self.num_inference_steps: Optional[int] = None
        self._reset_history()

    def _reset_history(self) -> None:
        self.model_outputs: List[Optional[torch.Tensor]] = [None] * self.solver_order
        self.timestep_list: List = [None] * self.solver_order
        self.lower_order_nums = 0
        self.last_sample: Optional[torch.Tensor] = None
        self.this_order = 0
        self._step_index: Optional[int] = None
        self._begin_index: Optional[int] = None
        #: per-step record of the integer decisions, for parity tests: (step_index, order, used_corrector)
        self.trace: List[Tuple[int, int, bool]] = []

    # -------------------------------------------------------------------------------- schedule
    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    def set_begin_index(self, begin_index: int = 0) -> None:
        self._begin_index = begin_index

    def set_timesteps(self, num_inference_steps: Optional[int] = None, device=None,
                      sigmas: Optional[Sequence[float]] = None, shift: Optional[float] = None) -> None:
        shift = self.shift if shift is None else shift
        if sigmas is None:
            if self.flavor == "twin":
                sigmas = np.linspace(self.sigma_max, self.sigma_min, num_inference_steps + 1).copy()[:-1]
                sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
            else:
                s = 1.0 - np.linspace(1, 1 / self.num_train_timesteps, num_inference_steps + 1)
                sigmas = np.flip(shift * s / (1 + (shift - 1) * s))[:-1].copy()
        else:
            sigmas = np.asarray(sigmas, dtype=np.float64)
            sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        timesteps = sigmas * self.num_train_timesteps                       # float64
        self.sigmas = torch.from_numpy(np.concatenate([sigmas, [0.0]]).astype(np.float32))  # stays on the host
        self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)   # truncation
        self.num_inference_steps = len(timesteps)
        self._reset_history()

    def index_for_timestep(self, timestep, schedule_timesteps: Optional[torch.Tensor] = None) -> int:
        ts = self.timesteps if schedule_timesteps is None else schedule_timesteps
        if isinstance(timestep, torch.Tensor):
            timestep = timestep.to(ts.device)
        hits = (ts == timestep).nonzero()
        # first `step` of a run takes the SECOND match when the value is duplicated (unipc.py:629-636)
        return hits[1 if len(hits) > 1 else 0].item()

    # -------------------------------------------------------------------------------- B(h) coefficients
    def _lambda(self, sigma: torch.Tensor) -> torch.Tensor:
        return torch.log(1 - sigma) - torch.log(sigma)

    def _bh_setup(self, sigma_t: torch.Tensor, sigma_s0: torch.Tensor, hist_sigmas: List[torch.Tensor], order: int):
        """Shared scalar algebra of UniP / UniC (unipc.py:401-452, 540-590); float32 0-dim host tensors."""
        lam_t, lam_s0 = self._lambda(sigma_t), self._lambda(sigma_s0)
        h = lam_t - lam_s0
        rks = [(self._lambda(s) - lam_s0) / h for s in hist_sigmas]
        rks_t = torch.tensor(rks + [1.0])
        hh = -h if self.predict_x0 else h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        B_h = hh if self.solver_type == "bh1" else torch.expm1(hh)
        R, b, fact = [], [], 1
        for i in range(1, order + 1):
            R.append(torch.pow(rks_t, i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        return rks, torch.stack(R), torch.tensor(b), h_phi_1, B_h

    def _predict(self, x: torch.Tensor, order: int) -> torch.Tensor:
        """UniP-order step from sigma[idx] to sigma[idx+1] (unipc.py:348-476)."""
        i = self._step_index
        m0 = self.model_outputs[-1]
        sigma_t, sigma_s0 = self.sigmas[i + 1], self.sigmas[i]
        hist = [self.sigmas[i - k] for k in range(1, order)]
        rks, R, b, h_phi_1, B_h = self._bh_setup(sigma_t, sigma_s0, hist, order)
        alpha_t = 1 - sigma_t
        D1s = [(self.model_outputs[-(k + 1)] - m0) / rks[k - 1] for k in range(1, order)]
        if D1s:
            rhos = torch.tensor([0.5], dtype=x.dtype) if order == 2 else torch.linalg.solve(R[:-1, :-1], b[:-1]).to(x.dtype)
        if self.predict_x0:
            x_t = sigma_t / sigma_s0 * x - alpha_t * h_phi_1 * m0
            lead = alpha_t * B_h
        else:
            x_t = alpha_t / (1 - sigma_s0) * x - sigma_t * h_phi_1 * m0
            lead = sigma_t * B_h
        if D1s:
            res = sum(float(r) * d for r, d in zip(rhos, D1s)) if len(D1s) > 1 else float(rhos[0]) * D1s[0]
            x_t = x_t - lead * res
        else:
            x_t = x_t - lead * 0
        return x_t.to(x.dtype)

    def _correct(self, model_t: torch.Tensor, last_sample: torch.Tensor, this_sample: torch.Tensor, order: int):
        """UniC-order correction of the sample at sigma[idx] using sigma[idx-1] history (unipc.py:478-622)."""
        i = self._step_index
        m0 = self.model_outputs[-1]
        x = last_sample
        sigma_t, sigma_s0 = self.sigmas[i], self.sigmas[i - 1]
        hist = [self.sigmas[i - (k + 1)] for k in range(1, order)]
        rks, R, b, h_phi_1, B_h = self._bh_setup(sigma_t, sigma_s0, hist, order)
        alpha_t = 1 - sigma_t
        D1s = [(self.model_outputs[-(k + 1)] - m0) / rks[k - 1] for k in range(1, order)]
        rhos = torch.tensor([0.5], dtype=x.dtype) if order == 1 else torch.linalg.solve(R, b).to(x.dtype)
        if self.predict_x0:
            x_t = sigma_t / sigma_s0 * x - alpha_t * h_phi_1 * m0
            lead = alpha_t * B_h
        else:
            x_t = alpha_t / (1 - sigma_s0) * x - sigma_t * h_phi_1 * m0
            lead = sigma_t * B_h
        corr = 0
        if D1s:
            corr = sum(float(r) * d for r, d in zip(rhos[:-1], D1s)) if len(D1s) > 1 else float(rhos[0]) * D1s[0]
        x_t = x_t - lead * (corr + rhos[-1] * (model_t - m0))
        return x_t.to(x.dtype)

    # -------------------------------------------------------------------------------- step
    def convert_model_output(self, model_output: torch.Tensor, sample: torch.Tensor) -> torch.Tensor:
        sigma_t = self.sigmas[self._step_index]
        if self.predict_x0:
            return sample - sigma_t * model_output          # x0 prediction (unipc.py:317-319)
        return sample - (1 - sigma_t) * model_output

    def step(self, model_output: torch.Tensor, timestep: Union[int, torch.Tensor], sample: torch.Tensor,
             return_dict: bool = True, generator=None):
        """unipc.py:651-737; ``return_dict`` defaults to True like the reference (a ``SchedulerOutput`` with ``.prev_sample``,
        also indexable); the engine passes ``return_dict=False`` and takes ``[0]`` (engine/wan/shared/__init__.py:569)."""
        if self.num_inference_steps is None:
            raise ValueError(
                "Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if self._step_index is None:
            self._step_index = self.index_for_timestep(timestep) if self._begin_index is None else self._begin_index
        idx = self._step_index
        use_corrector = idx > 0 and (idx - 1) not in self.disable_corrector and self.last_sample is not None
        converted = self.convert_model_output(model_output, sample)
        if use_corrector:
            sample = self._correct(converted, self.last_sample, sample, self.this_order)
        self.model_outputs = self.model_outputs[1:] + [converted]
        self.timestep_list = self.timestep_list[1:] + [timestep]
        this_order = min(self.solver_order, len(self.timesteps) - idx) if self.lower_order_final else self.solver_order
        self.this_order = min(this_order, self.lower_order_nums + 1)
        assert self.this_order > 0
        self.trace.append((idx, self.this_order, bool(use_corrector)))
        self.last_sample = sample
        prev = self._predict(sample, self.this_order)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        if return_dict:
            return SchedulerOutput(prev)
        return (prev,)

    def scale_model_input(self, sample: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        return sample


def get_timesteps(scheduler: UniPCMultistepScheduler, num_inference_steps: Optional[int] = None,
                  timesteps: Optional[Sequence[int]] = None, timesteps_as_indices: bool = False,
                  strength: float = 1.0, device=None) -> Tuple[torch.Tensor, int]:
    """BaseEngine._get_timesteps (apps/api/src/engine/base_engine.py:2971-3041) for this scheduler: default
    schedule, ``timesteps_as_indices`` gather ``scheduler.timesteps[num_train - ids]`` (:2989-3002) and the
    ``strength`` tail cut (:3032-3039)."""
    if timesteps is not None:
        if not timesteps_as_indices:
            raise ValueError(f"The current scheduler class {scheduler.__class__}'s `set_timesteps` does not support "
                             "custom timestep schedules. Please check whether you are using the correct scheduler.")
        ids = torch.tensor(list(timesteps), dtype=torch.long, device=scheduler.timesteps.device)
        ts = scheduler.timesteps[scheduler.num_train_timesteps - ids]
        scheduler.timesteps = ts
        scheduler.sigmas = scheduler.timesteps / scheduler.num_train_timesteps
        num_inference_steps = len(ts)
    else:
        scheduler.set_timesteps(num_inference_steps, device=device)
        ts = scheduler.timesteps
    if strength != 1.0:
        init = min(int(num_inference_steps * strength), num_inference_steps)
        t_start = max(num_inference_steps - init, 0)
        ts = ts[t_start * scheduler.order:]
        num_inference_steps = len(ts)
    return ts, num_inference_steps


class FlowMatchEulerDiscreteScheduler:
    """Flow-matching Euler scheduler the Flux manifests name (``diffusers.FlowMatchEulerDiscreteScheduler``,
    apps/api/manifest/image/flux-dev-text-to-image-1.0.0.v1.yml:38-46) with the FLUX.1-dev scheduler config
    (``use_dynamic_shifting=True``, exponential time shift).  The class lives in the un-vendored ``diffusers``
    dependency; this restates its published algorithm for the calls the Flux engine makes
    (engine/flux/t2i.py:110-135 ``set_timesteps(sigmas=linspace(1, 1/N, N), mu=...)``, engine/flux/shared.py:584-588
    ``step``) -- parity UNPINNED (no in-tree twin, no reference test).  ``step`` is two fp32 element-wise ops on the
    device (the latents of a 1024x1024 image are 262k elements)."""
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 3.0, use_dynamic_shifting: bool = True,
                 base_shift: float = 0.5, max_shift: float = 1.15, base_image_seq_len: int = 256,
                 max_image_seq_len: int = 4096):
        self.num_train_timesteps, self.shift, self.use_dynamic_shifting = num_train_timesteps, shift, use_dynamic_shifting
        self.base_shift, self.max_shift = base_shift, max_shift
        self.base_image_seq_len, self.max_image_seq_len = base_image_seq_len, max_image_seq_len
        self.config = self
        self.timesteps = self.sigmas = None
        self._step_index = self._begin_index = None

    def get(self, name, default=None):  # the engine reads scheduler.config.get("base_shift", 0.5) (t2i.py:121-127)
        return getattr(self, name, default)

    @property
    def step_index(self):
        return self._step_index

    def set_begin_index(self, begin_index: int = 0) -> None:
        self._begin_index = begin_index

    def set_timesteps(self, num_inference_steps: Optional[int] = None, device=None, sigmas=None, mu: Optional[float] = None):
        if self.use_dynamic_shifting and mu is None:
            raise ValueError("`mu` must be passed when `use_dynamic_shifting` is set to be `True`")
        if sigmas is None:
            ts = np.linspace(float(self.num_train_timesteps), 1.0, num_inference_steps)  # sigma_max=1, sigma_min=1/1000
            sigmas = ts / self.num_train_timesteps
        sigmas = np.array(sigmas).astype(np.float32)
        if self.use_dynamic_shifting:
            sigmas = math.exp(mu) / (math.exp(mu) + (1 / sigmas - 1) ** 1.0)
        else:
            sigmas = self.shift * sigmas / (1 + (self.shift - 1) * sigmas)
        sig = torch.from_numpy(np.asarray(sigmas)).to(dtype=torch.float32, device=device)
        self.timesteps = sig * self.num_train_timesteps
        self.sigmas = torch.cat([sig, torch.zeros(1, device=sig.device)])
        self._step_index = None
        return self.timesteps

    def index_for_timestep(self, timestep, schedule_timesteps: Optional[torch.Tensor] = None) -> int:
        """upstream ``FlowMatchEulerDiscreteScheduler.index_for_timestep``: equality lookup, SECOND match when duplicated."""
        ts = self.timesteps if schedule_timesteps is None else schedule_timesteps
        if isinstance(timestep, torch.Tensor):
            timestep = timestep.to(ts.device)
        hits = (ts == timestep).nonzero()
        if len(hits) == 0:
            raise ValueError(f"timestep {timestep} is not in the schedule")
        return hits[1 if len(hits) > 1 else 0].item()

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, return_dict: bool = True, **unused):
        if self._step_index is None:
            # upstream ``_init_step_index``: a strength- / img2img-truncated schedule without set_begin_index starts at the
            # position of ITS first timestep, not at sigma[0]
            self._step_index = self.index_for_timestep(timestep) if self._begin_index is None else self._begin_index
        i = self._step_index
        # upstream: ``dt = sigma_next - sigma`` (0-dim fp32 tensors); ``sample.float() + dt * model_output`` -- with a
        # bf16 model_output torch's type promotion makes the product a bf16 op (0-dim operands do not promote)
        dt = self.sigmas[i + 1] - self.sigmas[i]
        prev = sample.to(torch.float32) + dt * model_output
        self._step_index += 1
        prev = prev.to(model_output.dtype)
        return SchedulerOutput(prev) if return_dict else (prev,)


def calculate_shift(image_seq_len: int, base_seq_len: int = 256, max_seq_len: int = 4096, base_shift: float = 0.5,
                    max_shift: float = 1.15) -> float:
    """FluxShared.calculate_shift (engine/flux/shared.py:58-70): the ``mu`` of the dynamic time shift."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b
