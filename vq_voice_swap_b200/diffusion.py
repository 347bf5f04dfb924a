"""Continuous-time diffusion: noise schedules, the DDPM reverse step and the sampler loop.

Public surface of reference diffusion/{diffusion,schedule,make}.py.  The sampler
(`Diffusion.ddpm_sample`, reference diffusion.py:92-133) keeps the reference's call
signature and RNG call sequence (one randn_like per step except the last) so a
seeded run consumes the device generator identically; the per-step arithmetic of
`ddpm_previous` (:48-90) runs in libvqvs kernels -- fused into the UNet's final
store when the predictor is ours.
"""

import math
from abc import ABC, abstractmethod
from typing import Callable, Optional

import torch

from . import engine


class Schedule(ABC):
    @abstractmethod
    def __call__(self, t: torch.Tensor) -> torch.Tensor:
        """alpha-bar at continuous time t in [0, 1]."""


class ExpSchedule(Schedule):
    """alpha_bar(t) = exp(-k t^2) with alpha_bar(1) = alpha_final (reference schedule.py:15-31)."""

    def __init__(self, alpha_final: float = 1e-5):
        super().__init__()
        self.alpha_final = alpha_final
        self.k = -math.log(alpha_final)

    def __call__(self, t: torch.Tensor) -> torch.Tensor:
        return torch.exp(-self.k * (t ** 2))


class CosSchedule(Schedule):
    """alpha_bar(t) = cos(pi t / 2)^2 (reference schedule.py:34-41)."""

    def __call__(self, t: torch.Tensor) -> torch.Tensor:
        return torch.cos(t * math.pi / 2) ** 2


def make_schedule(name: str) -> Schedule:
    if name == "exp":
        return ExpSchedule()
    elif name == "cos":
        return CosSchedule()
    raise ValueError(f"unknown schedule: {name}")


def _per_sample(v: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    return v.to(like).reshape(-1, *([1] * (like.dim() - 1)))


class Diffusion:
    def __init__(self, schedule: Schedule):
        self.schedule = schedule

    # -- training-side helpers (plain torch; outside the sampling hot path) ----------------
    def sample_q(self, x_0, ts, epsilon=None):
        if epsilon is None:
            epsilon = torch.randn_like(x_0)
        a = _per_sample(self.schedule(ts), x_0)
        return a.sqrt() * x_0 + (1 - a).sqrt() * epsilon

    def eps_to_x0(self, x_t, ts, epsilon_prediction):
        a = _per_sample(self.schedule(ts), x_t)
        return (x_t - (1 - a).sqrt() * epsilon_prediction) * a.rsqrt()

    def x0_to_eps(self, x_t, ts, x_0):
        a = _per_sample(self.schedule(ts), x_t)
        return (x_t - x_0 * a.sqrt()) * (1 - a).rsqrt()

    def ddpm_losses(self, x, predictor, ts=None, noise=None):
        if ts is None:
            ts = torch.rand(len(x), device=x.device)
        if noise is None:
            noise = torch.randn_like(x)
        pred = predictor(self.sample_q(x, ts, epsilon=noise), ts)
        return ((noise - pred) ** 2).flatten(1).mean(dim=1)

    # -- reverse step -----------------------------------------------------------------------
    def step_coefficients(self, ts: torch.Tensor, step, sigma_large: bool = False):
        """Per-sample scalars of reference diffusion.py:64-78, evaluated with the same fp32 torch
        ops on the same device (the schedule is a user-visible callable on tensors)."""
        ab_t = self.schedule(ts)
        ab_p = self.schedule(ts - step)
        alpha = ab_t / ab_p
        beta = 1 - alpha
        sig2 = beta if sigma_large else beta * (1 - ab_p) / (1 - ab_t)
        one_m = 1 - ab_t
        packed = torch.stack(
            [alpha.rsqrt(), beta * one_m.rsqrt(), sig2.sqrt(), one_m.sqrt(), ab_t.rsqrt(), ab_t.sqrt(),
             one_m.rsqrt(), torch.zeros_like(ab_t)],
            dim=1,
        ).float().contiguous()
        return dict(packed=packed, alpha=alpha, beta=beta, sig2=sig2, ab_t=ab_t)

    def _update(self, x_t, ts, step, eps, noise, co, constrain, cond_fn, out=None):
        """x_{t-1} from an epsilon prediction held in a tensor (generic predictors / cond_fn)."""
        coef = co["packed"]
        if out is None:
            out = torch.empty_like(x_t)
        if cond_fn is not None:  # reference diffusion.py:80-83
            mean = engine.ddpm_finish(x_t, eps, coef, None, torch.empty_like(x_t))
            mean = mean + _per_sample(co["sig2"], x_t) * cond_fn(mean, ts - step)
            eps = (-mean * _per_sample(co["alpha"].sqrt(), x_t) + x_t) * _per_sample((1 - co["ab_t"]).sqrt(), x_t)
            eps = (eps / _per_sample(co["beta"], x_t)).contiguous()
        if not constrain:
            return engine.ddpm_finish(x_t, eps, coef, noise, out, None)
        # constrain subtracts the mean over the LAST axis only (reference diffusion.py:87: x0.mean(dim=-1, keepdim=True)),
        # i.e. per (sample, channel) row: run the row kernels over [N*C] rows of length T with the coefficients repeated
        n, t = x_t.shape[0], x_t.shape[-1]
        rows = x_t[0].numel() // t
        if rows > 1:
            coef = coef.repeat_interleave(rows, dim=0).contiguous()
        view = lambda a: None if a is None else a.reshape(n * rows, 1, t)  # noqa: E731
        x0_sum = engine.ddpm_x0_sum(view(x_t), view(eps), coef)
        engine.ddpm_finish(view(x_t), view(eps), coef, view(noise), view(out), x0_sum)
        return out

    def ddpm_previous(self, x_t, ts, step, epsilon_prediction, noise=None, sigma_large=False, constrain=False,
                      cond_fn: Callable = None):
        """Sample x_{t-step} given the epsilon prediction at t (reference diffusion.py:48-90)."""
        engine._require_cuda(x_t, epsilon_prediction)
        if noise is None:
            noise = torch.randn_like(x_t)
        with torch.no_grad():
            dtype = x_t.dtype
            x_t, eps = engine._f32(x_t), engine._f32(epsilon_prediction)
            ts = ts.to(x_t)
            co = self.step_coefficients(ts, step, sigma_large)
            return self._update(x_t, ts, step, eps, engine._f32(noise), co, constrain, cond_fn).to(dtype)

    # -- sampler ----------------------------------------------------------------------------
    def _step_tables(self, x_T, steps: int, sigma_large: bool, schedule: Optional[Callable]):
        """Every step's timestep and update coefficients, evaluated ONCE on the host (all rows of a batch share t, reference
        diffusion.py:115) with the same fp32 torch ops the reference runs per step on the device, then moved in one copy:
        the sampling loop itself launches no ATen kernel besides the noise draw."""
        n = x_T.shape[0]
        grid = torch.tensor([(i + 1) / steps for i in range(steps)][::-1], dtype=torch.float32)
        ts, t_step = grid, torch.full_like(grid, 1 / steps)
        if schedule is not None:  # sample-time remap (:116-118), applied elementwise like the reference does per step
            t_step = schedule(grid) - schedule(grid - 1 / steps)
            ts = schedule(grid)
        co = self.step_coefficients(ts, t_step, sigma_large)
        dev = x_T.device
        return dict(
            ts=ts.to(dev)[:, None].expand(steps, n).contiguous(),
            t_step=t_step.to(dev)[:, None].expand(steps, n).contiguous(),
            packed=co["packed"].to(dev)[:, None, :].expand(steps, n, 8).contiguous(),
            **{k: co[k].to(dev)[:, None].expand(steps, n).contiguous() for k in ("alpha", "beta", "sig2", "ab_t")},
        )

    def ddpm_sample(self, x_T, predictor, steps: int, progress: bool = False, sigma_large: bool = False,
                    constrain: bool = False, cond_fn: Callable = None, schedule: Callable = None,
                    noise_fn: Optional[Callable] = None):
        """Run `steps` reverse-diffusion steps from x_T (reference diffusion.py:92-133).

        noise_fn(step_index, like) -> noise tensor replaces the reference's `torch.randn_like(x_t)` draw of a step (an
        extension used by sharding.sample_sharded for noise keyed by the global sample index); by default the device
        generator is consumed exactly like the reference: one randn_like per step except the last."""
        engine._require_cuda(x_T)
        fast = engine.resolve_predictor(predictor)
        if fast is not None and fast[0].training and any(getattr(m, "dropout", 0.0) for m in fast[0].modules()):
            raise RuntimeError("dropout > 0 in train mode is not implemented on the sm_100a path: call model.eval()")
        x_t = engine._f32(x_T)
        bufs = [torch.empty_like(x_t), torch.empty_like(x_t)]
        with torch.no_grad():
            tab = self._step_tables(x_t, steps, sigma_large, schedule)
        its = range(steps)
        if progress:
            from tqdm.auto import tqdm

            its = tqdm(its)

        def draw(i):
            if i + 1 == steps:
                return None  # zeros on the last step (:127)
            return torch.randn_like(x_t) if noise_fn is None else engine._f32(noise_fn(i, x_t))

        for i in its:
            ts, t_step = tab["ts"][i], tab["t_step"][i]
            co = {k: tab[k][i] for k in ("packed", "alpha", "beta", "sig2", "ab_t")}
            with torch.no_grad():
                out = bufs[i & 1]
                if fast is not None:
                    # our predictor consumes no random numbers, so drawing the noise first keeps the generator sequence of
                    # the reference (which draws after the predictor call, :120-131) and lets the last kernel apply it
                    noise = draw(i)
                    net, cond, labels = fast
                    res = engine.fused_sample_step(net, x_t, ts, cond, labels, co["packed"], noise, out,
                                                   constrain, cond_fn is not None, first=(i == 0))
                    x_t = self._update(x_t, ts, t_step, res, noise, co, constrain, cond_fn, out) if cond_fn else res
                else:
                    eps = engine._f32(predictor(x_t, ts))
                    x_t = self._update(x_t, ts, t_step, eps, draw(i), co, constrain, cond_fn, out)
        return x_t.clone().to(x_T.dtype)
