"""log Z / ELBO / EUBO / LV / ESS from the log-weights (reference: BaseOCLoss.compute_results,
losses/oc.py:134-173; ESS, eval/metrics.py:134-140; evaluate_eubo, additions/hacking.py:24-32).

Each GPU reduces its own particles to 8 fp64 partials in the estimator kernel (lrds_estimator_partials):
    [0] m = max(-rnd)  [1] sum exp(-rnd - m)  [2] sum exp(2(-rnd - m))  [3] sum rnd  [4] sum rnd^2
    [5] count          [6] m' = max(rnd)      [7] sum exp(rnd - m')
Particles shard across GPUs with no data-path collective; the ONLY exchange of the whole path is one
all_gather of these 64 bytes per rank (NCCL over NVLink), followed by an associative host-side merge.
"""
from __future__ import annotations

import math

import torch

from . import _native as N


def estimator_partials(rnd: torch.Tensor, group=None) -> torch.Tensor:
    """8 fp64 partials of this process's log-weights (CUDA kernel), merged over ``group`` when given.
    Returns a CPU float64 tensor of shape (8,)."""
    if not rnd.is_cuda:
        raise N.LrdsError("estimator partials run on CUDA tensors only (no CPU fallback)")
    r = rnd.detach().reshape(-1).to(torch.float32).contiguous()
    B = r.numel()
    blocks = N.lib().lrds_estimator_blocks(B)
    scratch = torch.empty(8 * (blocks + 1), device=r.device, dtype=torch.float64)
    out = torch.empty(8, device=r.device, dtype=torch.float64)
    with torch.cuda.device(r.device):
        N.check(N.lib().lrds_estimator_partials(N.ptr(r), B, N.ptr(out), N.ptr(scratch), N.stream_ptr(r.device)))
    if group is not None:
        return gather_and_merge(out, group)
    return out.cpu()


def merge_partials(parts: torch.Tensor) -> torch.Tensor:
    """Associative merge of per-rank partial records (rows of ``parts``) in float64 on the host."""
    parts = parts.detach().to("cpu", torch.float64).reshape(-1, 8)
    parts = parts[parts[:, 5] > 0]
    if parts.shape[0] == 0:
        return torch.zeros(8, dtype=torch.float64)
    m = parts[:, 0].max()
    e = torch.exp(parts[:, 0] - m)
    m2 = parts[:, 6].max()
    out = torch.empty(8, dtype=torch.float64)
    out[0] = m
    out[1] = (parts[:, 1] * e).sum()
    out[2] = (parts[:, 2] * e * e).sum()
    out[3], out[4], out[5] = parts[:, 3].sum(), parts[:, 4].sum(), parts[:, 5].sum()
    out[6] = m2
    out[7] = (parts[:, 7] * torch.exp(parts[:, 6] - m2)).sum()
    return out


def gather_and_merge(partials: torch.Tensor, group=None, on_device: bool = False) -> torch.Tensor:
    """One all_gather of 8 doubles per rank (NCCL for CUDA tensors, gloo for CPU tensors) + the associative merge.
    CUDA tensors are merged by lrds_estimator_merge on the GPU; with ``on_device`` the merged record is returned
    without a host synchronisation (the caller reads it when it needs the scalars)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if partials.is_cuda:
        buf = torch.empty(world * 8, device=partials.device, dtype=torch.float64)
        dist.all_gather_into_tensor(buf, partials.contiguous(), group=group)
        out = torch.empty(8, device=partials.device, dtype=torch.float64)
        with torch.cuda.device(partials.device):
            N.check(N.lib().lrds_estimator_merge(N.ptr(buf), world, N.ptr(out), N.stream_ptr(partials.device)))
        return out if on_device else out.cpu()
    buf = [torch.empty_like(partials) for _ in range(world)]
    dist.all_gather(buf, partials.contiguous(), group=group)
    return merge_partials(torch.stack(buf))


def metrics_from_partials(p: torch.Tensor) -> dict:
    """Scalars of compute_results / get_metrics / evaluate_eubo from merged partials."""
    m, s1, s2, sr, sr2, n, m2, sf = (float(v) for v in p)
    out = {
        "elbo": -sr / n,
        "log_norm_const_is": m + math.log(s1) - math.log(n),
        "lv_loss": (sr2 - sr * sr / n) / (n - 1) if n > 1 else float("nan"),
        "effective_sample_size": s1 * s1 / s2,
        "log_norm_const_is_f": -(m2 + math.log(sf)) + math.log(n),
        "eubo": -sr / n,
        "count": n,
    }
    out["norm_effective_sample_size"] = out["effective_sample_size"] / n
    return out
