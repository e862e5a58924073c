"""GPU: the training objective through the fused rollout (sde_sampler_lrds_b200/train.py) - ``loss(ts, x, ...)`` with
method='lv' and the parameter gradient its ``backward()`` produces - against the CPU oracle's autograd and the
reference-generated gradient fixtures (tests/golden/grad_*.pt, oracle/make_golden.py --grads) on identical Brownian
increments; then one optimiser step through the solver API.

Tolerances: loss within 1e-4 relative; every parameter-gradient tensor within 1e-3 of its largest entry (2e-2 for the
logistic-regression target, whose clamp mask flips for ~1 % of the particles)."""
import math
import os

import pytest
import torch

from oracle import rollout_oracle as O
from tests.cases import GRAD_CASES, grad_case, initial_state, noise_for

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def worst(got: dict, want: dict):
    assert set(got) == set(want), set(got) ^ set(want)
    return max(((got[k] - want[k]).abs().max() / want[k].abs().max().clamp(min=1e-12)).item() for k in want)


@pytest.mark.parametrize("precision", ["fp32", "f16x3"])
@pytest.mark.parametrize("name", GRAD_CASES)
def test_lv_gradient_matches_oracle_and_reference(name, precision, device):
    from tests.product_builders import Built
    case = grad_case(name)
    gold = torch.load(os.path.join(GOLDEN, "grad_" + name + ".pt"))
    x0, noise = initial_state(case), noise_for(case)
    built = Built(case, device, precision)
    loss, metrics = built.train_loss(x0, noise)
    assert loss.requires_grad and loss.ndim == 0 and "train/n_filtered_cumulative" in metrics
    loss.backward()
    got = {n: p.grad.detach().cpu() for n, p in built.ctrl.named_parameters() if p.grad is not None}
    lo, go, _ = O.lv_loss_and_grads(case["problem"], x0, noise, max_rnd=None if case["problem"]["method"] == "cmcd" else 1e8,
                                    traj_per_sample=case.get("traj_per_sample", 1))
    tol = 2e-2 if case["problem"]["target"]["kind"] == "logreg" else 1e-3
    for want_loss, want, what in ((lo, go, "oracle"), (gold["loss"], gold["grads"], "reference")):
        assert abs(loss.item() - want_loss.item()) <= max(tol / 10, 1e-4) * max(1.0, abs(want_loss.item())), what
        assert worst(got, want) < tol, (what, worst(got, want))


def test_cached_plan_follows_parameter_updates(device):
    """A parameter update (an optimiser step) must not rebuild the O(K) host side of a rollout plan (tens of
    milliseconds of scalar schedule formulas) - but the next rollout must run with the new weights and TimeEmbed rows: the
    cached plan refreshed in place equals a freshly built one (to float32 rounding: the refresh evaluates the TimeEmbed
    rows on the device, a fresh plan on the host)."""
    from tests.cases import CASES
    from tests.product_builders import Built
    case = CASES["ei_many_modes"]()
    x0, noise = initial_state(case), noise_for(case)
    built = Built(case, device, "f16x3")
    _, r1, _ = built.simulate(x0, noise)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for p in built.ctrl.parameters():
            p.add_(0.02 * torch.randn(p.shape, generator=g).to(device))
    _, r2, _ = built.simulate(x0, noise)
    assert len(built.loss._plans) == 1
    fresh = Built(case, device, "f16x3")
    fresh.ctrl.load_state_dict(built.ctrl.state_dict())
    _, r3, _ = fresh.simulate(x0, noise)
    assert ((r2 - r3).abs() / r3.abs().clamp(min=1.0)).max() < 1e-5 and ((r1 - r2).abs() / r2.abs().clamp(min=1.0)).max() > 1e-3


@pytest.mark.parametrize("solver_type, kw", [
    ("vp-ref", dict(ref_type="gmm", integrator_type="ei", time_type="snr")),
    ("pis_orig", dict(ref_type="default", integrator_type="em", time_type="uniform")),
    ("dds_orig", dict(ref_type="default", integrator_type="em", time_type="uniform")),
    ("dis_orig", dict(ref_type="default", integrator_type="em", time_type="uniform")),
    ("dis_orig", dict(ref_type="default", integrator_type="em", time_type="uniform", model_type="target_informed_lerp_tempering")),
    ("cmcd", dict(ref_type="default", integrator_type="em", time_type="uniform")),
])
def test_training_steps_through_make_model(solver_type, kw, device):
    """make_model(...).step(): the LV loss is finite, every parameter of the control receives a gradient and moves, and
    the evaluation afterwards sees the updated weights (the packed weight images are rebuilt)."""
    from sde_sampler_lrds_b200 import benchmark_utils as BU
    from sde_sampler_lrds_b200.additions.hacking import TrainableWrapper
    from tests.test_solver_api_gpu import TRAIN, _gmm_ref, _randomise_last_layers
    d, M = 6, 5
    details = {"sigma": 1.0, **_gmm_ref(d, M)}
    kw = dict({"model_type": "target_informed_zero_init"}, **kw)
    model = BU.make_model(solver_type=solver_type, loss_type="lv", solver_details=details, target_details=BU.make_target_details("many_modes", dim=d, n_modes=M),
                          training_details=dict(TRAIN, train_steps=3), n_steps=24, device=str(device), **kw)
    _randomise_last_layers(model)
    before = {n: p.detach().clone() for n, p in model.generative_ctrl.named_parameters()}
    elbo0 = model.compute_results().metrics["eval/elbo"]
    res, hist = TrainableWrapper(model, verbose=False).run(keep_training_metrics=True)
    assert len(hist["train/loss"]) == 3 and all(math.isfinite(v) for v in hist["train/loss"])
    assert hist["train/no_grad"][-1] == 0 and hist["train/skipped_steps"][-1] == 0
    moved = [n for n, p in model.generative_ctrl.named_parameters() if not torch.equal(p.detach(), before[n])]
    assert len(moved) == len(before)
    assert math.isfinite(res.metrics["eval/elbo"]) and res.metrics["eval/elbo"] != elbo0
    # one plan for the training rollouts, one or two for the evaluation (with / without the Ito term): the cache key
    # must not depend on the identity of freshly bound methods
    assert len(model.loss._plans) <= 3, len(model.loss._plans)
    with pytest.raises(NotImplementedError):
        model.loss.method = "kl"
        model.compute_loss()


@pytest.mark.parametrize("name", ["ei_many_modes", "pis_many_modes", "cmcd_gmm"])
def test_training_with_generated_noise_equals_recorded_noise(name, device):
    """Without recorded increments the training rollout draws its noise in the kernel and the gradient pass regenerates
    it (lrds_normals): loss and gradient must equal the run that is handed those increments as a recorded array."""
    from sde_sampler_lrds_b200 import train as TR
    from tests.product_builders import Built
    case = grad_case(name)
    x0 = initial_state(case)
    K, B, d = noise_for(case).shape
    out = []
    for recorded in (False, True):
        built = Built(case, device, "f16x3")
        z = TR.normals(11, 0, K, B, d, device) if recorded else None
        loss, _ = built.train_loss(x0, z, seed=11)
        loss.backward()
        out.append((loss.item(), {n: p.grad.detach().clone() for n, p in built.ctrl.named_parameters() if p.grad is not None}))
    (l0, g0), (l1, g1) = out
    assert abs(l0 - l1) <= 1e-6 * max(1.0, abs(l1))
    # the two runs may pick different power-of-two scales for the fp16 cotangent operands of lrds_mlp_grad (an analytic
    # bound on the generator's normals vs. the maximum of the recorded array): equal up to that rounding noise
    assert worst(g0, g1) < 5e-4
