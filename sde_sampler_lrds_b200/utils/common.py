"""Host-side helpers with the reference's names: Results, get_timesteps, clip_and_log.

Mirrors sde_sampler/utils/common.py (Results 9-13, get_timesteps 30-82, clip_and_log 85-112).  Everything
here is O(K) time-grid work; nothing per-particle.
"""
from __future__ import annotations

import math
from collections import namedtuple

import torch

Results = namedtuple(
    "Results",
    "samples weights log_norm_const_preds expectation_preds ts xs metrics plots",
    defaults=[{}, {}, None, None, None, None, {}, {}],
)


def _bisect(fn, low, high, targets, iters):
    """Vectorised bisection for a decreasing ``fn`` (reference: binary_search_v, utils/common.py:18-27).  Like the
    reference, the first midpoint is whatever ``(low + high) / 2`` of the caller's scalars is (a Python float for float
    bounds: the SDE formulas then see a double-precision scalar, which is visible in the last bit of the grid)."""
    for _ in range(iters):
        mid = (low + high) / 2.0
        val = fn(mid)
        low = torch.where(val > targets, mid, low)
        high = torch.where(val <= targets, mid, high)
    return (low + high) / 2.0


def get_timesteps(start, end, dt=None, steps=None, rescale_t=None, n_attemps: int = 1024, sde=None, device=None):
    """Time grid of K+1 float32 knots: uniform, DDS cosine, or uniform in log-SNR of ``sde``.

    Same signature and results as the reference's get_timesteps (utils/common.py:30-82), including the
    spelling of ``n_attemps``.  The log-SNR grid is computed on the host copy of the SDE scalars."""
    if (steps is None) is (dt is None):
        raise ValueError("Exactly one of `dt` and `steps` should be defined.")
    if steps is None:
        steps = int(math.ceil((end - start) / dt))
    if sde is not None:
        host = sde.host() if hasattr(sde, "host") else sde
        if hasattr(sde, "host"):  # the scalar algebra runs on the host copy: tensor-valued bounds follow it
            start = start.cpu() if torch.is_tensor(start) else start
            end = end.cpu() if torch.is_tensor(end) else end
        snr0, snr1 = host.log_snr(start), host.log_snr(end)  # the caller's scalars as they are (utils/common.py:43-48)
        if torch.isnan(snr0):
            raise ValueError("NaN SNR at t_0")
        if torch.isnan(snr1):
            raise ValueError("NaN SNR at t_K")
        inner = torch.linspace(snr0, snr1, steps=steps + 1)[1:-1]
        knots = _bisect(host.log_snr, start, end, inner, n_attemps)
        grid = torch.concat([torch.FloatTensor([start]), knots, torch.FloatTensor([end])]).sort().values
        out_dev = device if device is not None else getattr(sde, "device", None)
        return grid.to(out_dev) if out_dev is not None else grid
    if rescale_t is None:
        return torch.linspace(start, end, steps=steps + 1, device=device)
    if rescale_t == "quad":
        end_t = torch.as_tensor(end)
        return torch.sqrt(torch.linspace(start, end_t.square(), steps=steps + 1, device=device)).clip(max=end)
    if rescale_t == "cosine":  # DDS discretisation (Vargas et al.), s = 0.008
        s = 0.008
        frac = torch.linspace(start, end, steps + 1, device=device) / end
        w = torch.cos(((frac + s) / (1 + s)) * torch.pi * 0.5) ** 4
        w = w / w.sum() * end
        return torch.concat((torch.tensor([start], device=device), torch.cumsum(w, -1)))
    raise ValueError("Unkown timestep rescaling method.")


def clip_and_log(tensor, max_norm=None, name=None, t=None, log_dt: float = 0.2):
    """Pure clamp to [-max_norm, max_norm] (the reference's logging branch is commented out)."""
    if max_norm is not None:
        tensor = tensor.clip(min=-1.0 * max_norm, max=max_norm)
    return tensor
