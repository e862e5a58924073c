"""The small public interfaces next to the rollout against fixtures produced by the unmodified reference
(oracle/make_golden.py --aux -> tests/golden/aux_grids_euler.pt): ``get_timesteps`` (utils/common.py:30-82) bit for bit
for every grid kind and SDE family the solvers build, ``EulerIntegrator.integrate`` (eq/integrator.py:84-129) with
recorded Brownian increments and off-grid output times."""
import os

import pytest
import torch

from tests.cases import AUX_GRIDS, AUX_SDES, euler_case
from tests.product_builders import build_sde

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "aux_grids_euler.pt"))


@pytest.mark.parametrize("name", list(AUX_GRIDS))
def test_get_timesteps_is_bit_equal_to_the_reference(name):
    from sde_sampler_lrds_b200.utils.common import get_timesteps
    sde_name, kw = AUX_GRIDS[name]
    sde = build_sde(AUX_SDES[sde_name], "cpu") if sde_name else None
    got, want = get_timesteps(sde=sde, **kw), GOLD["grids"][name]
    assert got.dtype == want.dtype == torch.float32 and got.shape == want.shape
    assert torch.equal(got, want), (got - want).abs().max().item()


@pytest.mark.parametrize("name", [n for n in AUX_GRIDS if AUX_GRIDS[n][0] in ("vp10", "vp20")])
def test_oracle_get_timesteps_within_one_ulp(name):
    """The oracle's grid (tensor-valued first bisection step) may differ from the reference's in the last bit."""
    from oracle import rollout_oracle as O
    sde_name, kw = AUX_GRIDS[name]
    got = O.get_timesteps(sde=O.make_sde(AUX_SDES[sde_name]), **kw)
    want = GOLD["grids"][name]
    assert got.shape == want.shape and (got - want).abs().max() <= 2.4e-7


@pytest.mark.gpu
def test_get_timesteps_of_a_device_resident_sde(device):
    from sde_sampler_lrds_b200.utils.common import get_timesteps
    for name in ("snr_vp20", "snr_pbm", "snr_vpcos"):
        sde_name, kw = AUX_GRIDS[name]
        got = get_timesteps(sde=build_sde(AUX_SDES[sde_name], device), **kw)
        assert got.device.type == "cuda" and torch.equal(got.cpu(), GOLD["grids"][name])


@pytest.mark.gpu
@pytest.mark.parametrize("snr", [False, True])
def test_euler_integrator_matches_the_reference_class(device, snr):
    from sde_sampler_lrds_b200.eq.integrator import EulerIntegrator
    ec = euler_case()
    sde = build_sde(ec["sde"], device)
    it = iter(ec["noise"].to(device))

    def bm(s, t):
        return next(it) * torch.sqrt(t - s)
    if snr:
        got = EulerIntegrator(dt=None, steps=ec["snr_steps"]).integrate(sde, ec["ts_snr"].to(device), ec["x0"].to(device), bm=bm,
                                                                         snr_adapted=True)
        want = GOLD["euler_xs_snr"]
    else:
        got = EulerIntegrator(dt=ec["dt"]).integrate(sde, ec["ts"].to(device), ec["x0"].to(device), bm=bm)
        want = GOLD["euler_xs"]
    assert got.shape == want.shape
    err = ((got.cpu() - want).abs() / want.abs().clamp(min=1.0)).max().item()
    assert err < 2e-6, err


@pytest.mark.gpu
def test_euler_integrator_own_noise_has_the_right_law(device):
    """Without an injected Brownian motion the increments come from lrds_normals: over a driftless unit-diffusion SDE
    the end point is N(x0, T)."""
    from sde_sampler_lrds_b200.eq.integrator import EulerIntegrator

    class BM:
        def drift(self, t, x):
            return torch.zeros_like(x)

        def diff(self, t, x):
            return torch.ones((), device=x.device)
    x0 = torch.zeros(20000, 4, device=device)
    ts = torch.tensor([0.0, 2.0], device=device)
    xs = EulerIntegrator(dt=0.1).integrate(BM(), ts, x0)
    assert xs.shape == (2, 20000, 4)
    end = xs[-1]
    assert abs(end.mean().item()) < 0.05 and abs(end.var().item() - 2.0) < 0.08
