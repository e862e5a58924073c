"""MALA and random-walk Metropolis chains (lrds_mala; sde_sampler/additions/mcmc.py: mala_step 75-134, rwmh_step 258-290
+ the mcmc_sample loop of experiments/benchmark_utils.py).

CPU: the oracle's restatement against outputs of the reference's own mala_step / heuristics_step_size
(tests/golden/mala_*.pt, oracle/make_golden.py --mala) on identical draws.  GPU: the one-launch kernel against both.
A Metropolis test is a discontinuity: a chain whose accept decision flips (log u within float rounding of the log
acceptance ratio) follows a different path from there on, so per-chain agreement is required for >= 90 % of the chains
(measured: 100 %), each of them within 1e-4 relative over the whole recorded path."""
import os

import pytest
import torch

from oracle import rollout_oracle as O
from tests.cases import MALA_CASES, mala_inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def oracle_chains(case, *a, **k):
    return (O.rwmh_chains if case.get("mcmc_type") == "rwmh" else O.mala_chains)(*a, **k)


def chains_within(got, want, tol=1e-4):
    """fraction of chains whose whole path ys[:, c, :] agrees within tol relative (denominator max(|want|, 1))."""
    err = ((got - want).abs() / want.abs().clamp(min=1.0)).amax(dim=(0, 2))
    return (err <= tol).float().mean().item(), err.max().item()


@pytest.mark.parametrize("name", list(MALA_CASES))
def test_oracle_mala_matches_reference_golden(name):
    case = MALA_CASES[name]()
    gold = torch.load(os.path.join(GOLDEN, name + ".pt"))
    y_init, noise, unif = mala_inputs(case)
    ys, h, acc = oracle_chains(case, case["target"], y_init, case["step_size"], case["n_warmup"], case["n_steps"], noise, unif)
    assert ys.shape == gold["ys"].shape
    f, worst = chains_within(ys, gold["ys"])
    assert f >= 0.9, (f, worst)
    assert ((h - gold["step_size"]).abs() <= 1e-6 * gold["step_size"]).float().mean() >= 0.9


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(MALA_CASES))
def test_mala_kernel_matches_oracle_and_reference(name, device):
    from sde_sampler_lrds_b200.additions.mcmc import mala_chains
    from tests.product_builders import build_target
    case = MALA_CASES[name]()
    gold = torch.load(os.path.join(GOLDEN, name + ".pt"))
    y_init, noise, unif = mala_inputs(case)
    target = build_target(case["target"], device)
    ys, h, acc = mala_chains(target, y_init.to(device), case["step_size"], case["n_warmup"], case["n_steps"],
                             noise=noise, unif=unif, return_log_acc=True, mcmc_type=case.get("mcmc_type", "mala"))
    yo, ho, ao = oracle_chains(case, case["target"], y_init, case["step_size"], case["n_warmup"], case["n_steps"], noise, unif)
    assert ys.shape == gold["ys"].shape and h.shape == (y_init.shape[0], 1) and torch.isfinite(ys).all()
    for want, what in ((yo, "oracle"), (gold["ys"], "reference")):
        f, worst = chains_within(ys.cpu(), want)
        assert f >= 0.9, (what, f, worst)
    ok = ((h.cpu() - gold["step_size"]).abs() <= 1e-5 * gold["step_size"]).float().mean().item()
    assert ok >= 0.9, ok
    # the first step's log acceptance ratio has no history behind it: every chain must agree
    assert ((acc[0].cpu() - gold["log_acc"][0]).abs() <= 1e-3 * gold["log_acc"][0].abs().clamp(min=1.0)).all()


@pytest.mark.gpu
def test_mala_in_kernel_draws_follow_the_philox_spec(device):
    """Production mode (in-kernel Philox normals and uniforms) equals validation mode fed with oracle/philox_ref.py's
    statement of the same streams, up to the SFU approximations of the normals (chains may flip: >= 90 %)."""
    from sde_sampler_lrds_b200.additions.mcmc import mala_chains
    from tests.product_builders import build_target
    from oracle import philox_ref
    case = MALA_CASES["mala_many_modes"]()
    y_init, _, _ = mala_inputs(case)
    C, d = y_init.shape
    S = case["n_warmup"] + case["n_steps"]
    target = build_target(case["target"], device)
    seed = 0x1234_5678_9ABC
    a, _ = mala_chains(target, y_init.to(device), case["step_size"], case["n_warmup"], case["n_steps"], seed=seed)
    noise = torch.from_numpy(philox_ref.normals(seed, C, S, d))
    unif = torch.from_numpy(philox_ref.uniforms(seed, C, S))
    b, _ = mala_chains(target, y_init.to(device), case["step_size"], case["n_warmup"], case["n_steps"], noise=noise, unif=unif)
    f, worst = chains_within(a.cpu(), b.cpu(), tol=1e-3)
    assert f >= 0.9, (f, worst)


@pytest.mark.gpu
def test_mcmc_sample_api(device):
    """mcmc_sample with the reference's signature: dataset of dataset_length rows on the CPU, chains that stay near
    their modes, then a GMM reference fitted to it could be handed to RDS.change_reference_type."""
    from sde_sampler_lrds_b200 import benchmark_utils as BU
    from tests.product_builders import build_target
    case = MALA_CASES["mala_many_modes"]()
    target = build_target(case["target"], device)
    data = BU.mcmc_sample(device, target, case["x_init"], step_size=0.05, n_chains_per_mode=4, dataset_length=2800,
                          n_warmup_steps=64, seed=5)
    assert data.shape == (2800, 10) and data.device.type == "cpu" and torch.isfinite(data).all()
    # every sample sits within a few standard deviations of one of the modes
    dist = torch.cdist(data, case["target"]["loc"]).min(dim=1).values
    assert dist.max() < 6.0 * (0.5 * 10) ** 0.5
    rw = BU.mcmc_sample(device, target, case["x_init"], mcmc_type="rwmh", step_size=0.2, n_chains_per_mode=4,
                        dataset_length=1400, n_warmup_steps=64, seed=6)
    assert rw.shape == (1400, 10) and torch.isfinite(rw).all()
    assert torch.cdist(rw, case["target"]["loc"]).min(dim=1).values.max() < 6.0 * (0.5 * 10) ** 0.5
    with pytest.raises(NotImplementedError):
        BU.mcmc_sample(device, target, case["x_init"], mcmc_type="ula")
    # the reference-fitting workflow of the experiments: MCMC data -> diagonal GMM by EM -> RDS with that reference
    weights, means, variances = BU.fit_gmm(7, data, means_init=case["x_init"])
    assert weights.shape == (7,) and means.shape == (7, 10) and variances.shape == (7, 10) and (variances > 0).all()
    assert torch.cdist(means, case["target"]["loc"]).min(dim=1).values.max() < 1.0
    details = {"sigma": 1.0, "weights_ref": weights, "means_ref": means, "variances_ref": variances}
    model = BU.make_model(solver_type="vp-ref", ref_type="gmm", loss_type="lv", integrator_type="ei",
                          model_type="target_informed_zero_init", time_type="snr", solver_details=details,
                          target_details=BU.make_target_details("many_modes", dim=10, n_modes=7),
                          training_details={"train_steps": 2, "train_batch_size": 64, "eval_batch_size": 512},
                          n_steps=32, device=str(device))
    res = model.compute_results()
    # a reference fitted to the target itself with a zero-initialised control is already a good sampler
    assert res.metrics["eval/elbo"] > -5.0 and abs(res.log_norm_const_preds["log_norm_const_is"]) < 1.0
