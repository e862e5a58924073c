"""phi^4 lattice target (reference: sde_sampler/distr/phi_four.py; 1-D lattice, Dirichlet-0 boundary, no
tilt - the only configuration whose closed-form grad_U the reference implements, lines 81-90)."""
from __future__ import annotations

import torch

from .. import _native as N
from .base import Distribution


class PhiFour(Distribution):
    def __init__(self, a, b, dim, dim_phys=1, beta=1, bc=("dirichlet", 0), tilt=None, grid_points=1024, **kwargs):
        if dim_phys != 1 or tuple(bc) != ("dirichlet", 0) or tilt is not None:
            raise NotImplementedError("PhiFour kernel covers dim_phys=1, Dirichlet-0, no tilt (as grad_U in the reference)")
        self.a, self.b, self.beta = a, b, beta
        self.dim_grid, self.dim_phys = dim, dim_phys
        self.bc, self.tilt = bc, tilt
        self.coef = self.a * self.dim_grid
        super().__init__(dim=dim, grid_points=grid_points, **kwargs)
        self.set_domain(torch.stack([-1.5 * torch.ones((dim,)), 1.5 * torch.ones((dim,))], dim=1))

    def _lrds_pack(self, device):
        d = N.Distr()
        d.kind = N.DISTR_PHI4
        d.phi4.a, d.phi4.b, d.phi4.beta = float(self.a), float(self.b), float(self.beta)
        return d, None

    def U(self, x):
        return -self.unnorm_log_prob(x).squeeze(-1) / self.beta

    def grad_U(self, x):
        return -self.score(x) / self.beta

    def compute_phi_four_weight(self, samples):
        mask = samples[:, int(self.dim / 2)] > 0
        return (1.0 - mask.float().mean()) / mask.float().mean()
