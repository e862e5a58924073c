"""Dirac prior used by PIS / pinned-BM solvers (reference: sde_sampler/distr/delta.py:8-31)."""
from __future__ import annotations

import torch

from .gauss import Gauss


class Delta(Gauss):
    def __init__(self, dim: int = 1, loc=0.0, approx_scale: float = 1e-3, domain_scale: float = 10, **kwargs):
        super().__init__(dim=dim, loc=loc, scale=approx_scale, domain_scale=domain_scale, **kwargs)

    def sample(self, shape=None) -> torch.Tensor:
        return self.loc.repeat(*tuple(shape or ()), 1)
