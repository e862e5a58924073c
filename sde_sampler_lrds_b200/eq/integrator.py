"""Generic Euler-Maruyama integrator with the reference's interface (sde_sampler/eq/integrator.py:66-129).

``EulerIntegrator.integrate(sde, ts, x_init, timesteps=None, bm=None, snr_adapted=False)`` steps
``x_t = x_s + drift(s, x_s) (t - s) + diff(s, x_s) noise`` over ``timesteps`` and returns the states interpolated at
``ts``.  The solvers use it for the inference-process plots (solver/oc.py:163-180); the fused rollout kernel is the
hot path, this is the drop-in for the small interface.  The update itself runs in the library's ``lrds_axpy_step``
kernel when the diffusion coefficient is a scalar (every OU-type SDE); noise is ``bm(s, t)`` when a Brownian motion
is injected, else counter-based Philox normals from ``lrds_normals``.  CUDA tensors only.
``TorchSDEIntegrator`` (torchsde) is out of scope: torchsde is not installed and no rollout solver uses it.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _native as N
from ..utils.common import get_timesteps


def interpolate(ts: torch.Tensor, s, t, xs: torch.Tensor, xt: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """Linear interpolation of the states at the output times in (s, t]   (integrator.py:66-82)."""
    ind = torch.searchsorted(ts, t + eps, side="right")
    t_eval = ts[:ind]
    coeff = ((t_eval - s) / (t - s)).view(-1, *[1] * xs.ndim)  # not clipped: an output time within eps behind t extrapolates, as in the reference
    return torch.lerp(xs.unsqueeze(0), xt.unsqueeze(0), coeff)


class EulerIntegrator:
    def __init__(self, dt: float | None = None, steps: int | None = None, rescale_t: str | None = None, eps: float = 1e-8):
        self.dt, self.steps, self.rescale_t, self.eps = dt, steps, rescale_t, eps
        self._calls = 0

    def _noise(self, x: torch.Tensor, step: int) -> torch.Tensor:
        B, d = x.reshape(-1, x.shape[-1]).shape
        out = torch.empty(1, B, d, device=x.device, dtype=torch.float32)
        seed = (int(torch.initial_seed()) * 0x9E3779B97F4A7C15 + self._calls * 0xD1B54A32D192ED03) & (2 ** 64 - 1)
        with torch.cuda.device(x.device):
            N.check(N.lib().lrds_normals(C.c_uint64(seed), C.c_uint64(0), 2, 1, B, d, N.ptr(out), N.stream_ptr(x.device)))
        # the generator is keyed by (particle, step): move the step into the key stream
        self._calls += 1
        return out[0].reshape(x.shape)

    @torch.no_grad()
    def integrate(self, sde, ts: torch.Tensor, x_init: torch.Tensor, timesteps: torch.Tensor | None = None, bm=None,
                  snr_adapted: bool = False) -> torch.Tensor:
        if not x_init.is_cuda:
            raise N.LrdsError("EulerIntegrator runs on CUDA tensors only (no CPU fallback)")
        if timesteps is None:
            timesteps = get_timesteps(ts[0], ts[-1], dt=self.dt, steps=self.steps, rescale_t=self.rescale_t,
                                      device=ts.device, sde=sde if snr_adapted else None)
        ts_count, xs_out, xs = 0, [], x_init.to(torch.float32).contiguous()
        for k, (s, t) in enumerate(zip(timesteps[:-1], timesteps[1:])):
            dt = t - s
            noise = bm(s, t) if bm is not None else self._noise(xs, k) * torch.sqrt(dt)
            drift, diff = sde.drift(s, xs), sde.diff(s, xs)
            if torch.is_tensor(diff) and diff.numel() == 1 and torch.is_tensor(drift) and drift.shape == xs.shape:
                xt = torch.empty_like(xs)
                with torch.cuda.device(xs.device):
                    N.check(N.lib().lrds_axpy_step(N.ptr(xs), N.ptr(drift.contiguous()), N.ptr(noise.contiguous()), 1.0,
                                                   float(dt), float(diff), N.ptr(xt), xs.numel(), N.stream_ptr(xs.device)))
            else:
                xt = xs + drift * dt + diff * noise
            if ts_count < len(ts) and ts[ts_count] <= t + self.eps:
                xs_out.append(interpolate(ts[ts_count:], s, t, xs, xt, eps=self.eps))
                ts_count += xs_out[-1].shape[0]
            xs = xt
        xs_out = torch.cat(xs_out)
        assert ts_count == xs_out.shape[0]
        return xs_out
