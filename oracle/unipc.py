"""CPU ORACLE (test infrastructure) for the scheduler side of the denoise loop -- numpy restatement of the
reference's flow-UniPC scheduler and of the loop's integer decisions.

Follows /root/reference/apps/api/src/scheduler/unipc.py (in-repo twin of the un-vendored
``diffusers.UniPCMultistepScheduler`` that the Wan manifests name):
    :112-131   training sigma schedule (float32), sigma_min / sigma_max
    :159-227   set_timesteps  -> sigmas (float64 -> float32), timesteps = trunc(sigma*1000) as int64
    :624-636   index_for_timestep ("second match" rule)
    :651-737   step: corrector gate, history shift, order warm-up, predictor, step_index += 1
    :348-476   multistep_uni_p_bh_update      :478-622   multistep_uni_c_bh_update
and engine/wan/shared/__init__.py:335-337,464-476 (expert / guidance selection by ``t >= boundary``),
engine/base_engine.py:2989-3002,3032-3039 (timesteps_as_indices gather, strength cut).

Pinned by tests/golden/unipc_*.npz, produced by running the reference class itself (oracle/make_golden.py):
integer outputs must match exactly, float outputs to 2e-6 relative (numpy vs torch libm differences).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32


def train_sigmas(num_train_timesteps: int = 1000, shift: float = 1.0) -> np.ndarray:
    alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
    s = (1.0 - alphas).astype(F32)
    return (F32(shift) * s / (F32(1) + F32(shift - 1) * s)).astype(F32)


def make_schedule(num_inference_steps: int, shift: float = 1.0, num_train_timesteps: int = 1000,
                  flavor: str = "twin") -> Tuple[np.ndarray, np.ndarray]:
    """-> (sigmas float32 [n+1] ending in 0, timesteps int64 [n])."""
    if flavor == "twin":
        tr = train_sigmas(num_train_timesteps, shift)
        sigma_max, sigma_min = float(tr[0]), float(tr[-1])
        sig = np.linspace(sigma_max, sigma_min, num_inference_steps + 1).copy()[:-1]
        sig = shift * sig / (1 + (shift - 1) * sig)
    elif flavor == "diffusers":
        s = 1.0 - np.linspace(1, 1 / num_train_timesteps, num_inference_steps + 1)
        sig = np.flip(shift * s / (1 + (shift - 1) * s))[:-1].copy()
    else:
        raise ValueError(flavor)
    timesteps = (sig * num_train_timesteps).astype(np.int64)  # float64 -> int64 truncation
    sigmas = np.concatenate([sig, [0.0]]).astype(F32)
    return sigmas, timesteps


def index_for_timestep(timesteps: np.ndarray, t: int) -> int:
    hits = np.nonzero(timesteps == t)[0]
    return int(hits[1 if len(hits) > 1 else 0])


def step_orders(num_steps: int, solver_order: int = 2, lower_order_final: bool = True,
                disable_corrector: Sequence[int] = ()) -> List[Tuple[int, int, bool]]:
    """Integer state machine of `step` for a full run from index 0: [(step_index, this_order, use_corrector)]."""
    out, lower, have_last = [], 0, False
    for idx in range(num_steps):
        use_corr = idx > 0 and (idx - 1) not in disable_corrector and have_last
        order = min(solver_order, num_steps - idx) if lower_order_final else solver_order
        order = min(order, lower + 1)
        out.append((idx, order, use_corr))
        have_last = True
        if lower < solver_order:
            lower += 1
    return out


def expert_and_guidance(timesteps: np.ndarray, boundary_timestep: Optional[float], guidance_scale):
    """Per step: ("high"|"low", guidance) -- t >= boundary picks the high-noise expert and guidance_scale[0]."""
    res = []
    for t in timesteps:
        high = boundary_timestep is not None and t >= boundary_timestep
        g = float(guidance_scale[0] if high else guidance_scale[1]) if isinstance(guidance_scale, (list, tuple)) \
            else float(guidance_scale)
        res.append(("high" if high else "low", g))
    return res


def timesteps_as_indices(schedule_timesteps: np.ndarray, ids: Sequence[int], num_train_timesteps: int = 1000):
    """base_engine.py:2989-3002: timesteps = scheduler.timesteps[num_train - ids]."""
    return schedule_timesteps[num_train_timesteps - np.asarray(ids, dtype=np.int64)]


def strength_cut(timesteps: np.ndarray, strength: float, order: int = 1) -> np.ndarray:
    n = len(timesteps)
    init = min(int(n * strength), n)
    return timesteps[max(n - init, 0) * order:]


# --------------------------------------------------------------------------------------------------
# float side: B(h) predictor / corrector in float32 (solver_type bh2 / bh1, predict_x0)
# --------------------------------------------------------------------------------------------------
def _lam(s):
    with np.errstate(divide="ignore"):
        return (np.log(F32(1) - s) - np.log(s)).astype(F32)


def _setup(sigma_t, sigma_s0, hist, order, solver_type):
    with np.errstate(divide="ignore", invalid="ignore"):
        lam_t, lam_s0 = _lam(sigma_t), _lam(sigma_s0)
        h = F32(lam_t - lam_s0)
        rks = [F32((_lam(s) - lam_s0) / h) for s in hist]
        rks_t = np.array(rks + [F32(1.0)], dtype=F32)
        hh = F32(-h)
        h_phi_1 = F32(np.expm1(hh))
        h_phi_k = F32(h_phi_1 / hh - F32(1))
        B_h = hh if solver_type == "bh1" else F32(np.expm1(hh))
        R, b, fact = [], [], 1
        for i in range(1, order + 1):
            R.append(np.power(rks_t, F32(i - 1)).astype(F32))
            b.append(F32(h_phi_k * F32(fact) / B_h))
            fact *= i + 1
            h_phi_k = F32(h_phi_k / hh - F32(1 / fact))
    return rks, np.stack(R), np.array(b, dtype=F32), h_phi_1, B_h


class UniPCOracle:
    """Stateful float32 restatement (predict_x0=True, flow_prediction)."""

    def __init__(self, sigmas: np.ndarray, timesteps: np.ndarray, solver_order: int = 2, solver_type: str = "bh2",
                 lower_order_final: bool = True, disable_corrector: Sequence[int] = ()):
        self.sigmas, self.timesteps = sigmas.astype(F32), timesteps
        self.solver_order, self.solver_type = solver_order, solver_type
        self.lower_order_final, self.disable_corrector = lower_order_final, list(disable_corrector)
        self.model_outputs: List[Optional[np.ndarray]] = [None] * solver_order
        self.lower_order_nums, self.last_sample, self.step_index, self.this_order = 0, None, None, 0
        self.trace: List[Tuple[int, int, bool]] = []

    def _update(self, x, m0, sigma_t, sigma_s0, hist, order, model_t=None):
        rks, R, b, h_phi_1, B_h = _setup(sigma_t, sigma_s0, hist, order, self.solver_type)
        alpha_t = F32(1) - sigma_t
        D1s = [((self.model_outputs[-(k + 1)] - m0) / rks[k - 1]).astype(F32) for k in range(1, order)]
        x_t = (F32(sigma_t / sigma_s0) * x - F32(alpha_t * h_phi_1) * m0).astype(F32)
        lead = F32(alpha_t * B_h)
        if model_t is None:  # predictor
            if D1s:
                rhos = np.array([0.5], dtype=F32) if order == 2 else np.linalg.solve(R[:-1, :-1], b[:-1]).astype(F32)
                res = sum(r * d for r, d in zip(rhos, D1s))
                x_t = (x_t - lead * res).astype(F32)
            return x_t
        rhos = np.array([0.5], dtype=F32) if order == 1 else np.linalg.solve(R, b).astype(F32)
        corr = sum(r * d for r, d in zip(rhos[:-1], D1s)) if D1s else F32(0)
        return (x_t - lead * (corr + rhos[-1] * (model_t - m0))).astype(F32)

    def step(self, model_output: np.ndarray, t: int, sample: np.ndarray) -> np.ndarray:
        if self.step_index is None:
            self.step_index = index_for_timestep(self.timesteps, t)
        i = self.step_index
        use_corr = i > 0 and (i - 1) not in self.disable_corrector and self.last_sample is not None
        converted = (sample - self.sigmas[i] * model_output).astype(F32)
        if use_corr:
            hist = [self.sigmas[i - (k + 1)] for k in range(1, self.this_order)]
            sample = self._update(self.last_sample, self.model_outputs[-1], self.sigmas[i], self.sigmas[i - 1], hist,
                                  self.this_order, model_t=converted)
        self.model_outputs = self.model_outputs[1:] + [converted]
        order = min(self.solver_order, len(self.timesteps) - i) if self.lower_order_final else self.solver_order
        self.this_order = min(order, self.lower_order_nums + 1)
        self.trace.append((i, self.this_order, bool(use_corr)))
        self.last_sample = sample
        hist = [self.sigmas[i - k] for k in range(1, self.this_order)]
        prev = self._update(sample, converted, self.sigmas[i + 1], self.sigmas[i], hist, self.this_order)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self.step_index += 1
        return prev
