"""GPU: the reference-level entry points above the loss objects - make_model, Solver.compute_results,
TrainableWrapper.evaluate / evaluate_eubo, EulerIntegrator - run through the fused kernels and agree with the
oracle's estimators on the log-weights they produce."""
import math

import pytest
import torch

from oracle import rollout_oracle as O

pytestmark = pytest.mark.gpu
TRAIN = {"train_steps": 16, "train_batch_size": 64, "eval_batch_size": 1000}


def _gmm_ref(d, M, seed=3):
    g = torch.Generator().manual_seed(seed)
    return {"weights_ref": torch.rand(M, generator=g) + 0.5, "means_ref": torch.randn(M, d, generator=g) * 2.0,
            "variances_ref": torch.rand(M, d, generator=g) + 0.5}


def _randomise_last_layers(model, gain=0.5):
    """The reference zero-inits the control (scale 1e-6); give it O(1) outputs so that the test is not vacuous."""
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for m in (model.generative_ctrl.base_model.out_layer, getattr(model.generative_ctrl, "score_model", None) and
                  model.generative_ctrl.score_model.out_layer):
            if m:
                m.weight.copy_((torch.rand(m.weight.shape, generator=g) * 2 - 1) * gain / math.sqrt(m.weight.shape[1]))
                m.bias.copy_((torch.rand(m.bias.shape, generator=g) * 2 - 1) * 0.1 + 0.2)


@pytest.mark.parametrize("solver_type, kw", [
    ("vp-ref", dict(ref_type="gmm", integrator_type="ei", time_type="snr")),
    ("vp-ref", dict(ref_type="gaussian", integrator_type="ddpm_like", time_type="snr")),
    ("vp-ref", dict(ref_type="default", integrator_type="em", time_type="uniform", model_type="base_zero_init")),
    ("vp-ref", dict(ref_type="gmm", integrator_type="ei", time_type="uniform", force_vp_cosine=True)),
    ("pbm-ref", dict(ref_type="default", integrator_type="ei", time_type="snr")),
    ("pis_orig", dict(ref_type="default", integrator_type="em", time_type="uniform")),
    ("dds_orig", dict(ref_type="default", integrator_type="em", time_type="uniform")),
    ("cmcd", dict(ref_type="gaussian", integrator_type="em", time_type="uniform")),
    ("dis_orig", dict(ref_type="default", integrator_type="em", time_type="uniform")),
    ("dis_orig", dict(ref_type="default", integrator_type="em", time_type="uniform", force_vp20=True)),
    ("dis_orig", dict(ref_type="default", integrator_type="em", time_type="uniform", model_type="target_informed_lerp_tempering")),
    ("dis_orig", dict(ref_type="default", integrator_type="em", time_type="uniform", model_type="target_informed_langevin_init")),
])
def test_make_model_evaluate(solver_type, kw, device):
    from sde_sampler_lrds_b200 import benchmark_utils as BU
    from sde_sampler_lrds_b200.additions.hacking import TrainableWrapper
    d, M = 6, 5
    details = {"sigma": 1.0, **_gmm_ref(d, M), "mean_ref": torch.zeros(d), "var_ref": torch.full((d,), 2.0),
               "mean": torch.zeros(d), "var": torch.full((d,), 4.0)}
    args = dict(solver_type=solver_type, loss_type="lv", model_type="target_informed_zero_init",
                solver_details=details, target_details=BU.make_target_details("many_modes", dim=d, n_modes=M),
                training_details=TRAIN, n_steps=24, device=str(device))
    args.update(kw)
    model = BU.make_model(**args)
    _randomise_last_layers(model)
    res = TrainableWrapper(model, verbose=False).evaluate()
    B = TRAIN["eval_batch_size"]
    K1 = len(model.eval_ts)
    assert res.samples.shape == (B, d) and res.weights.shape == (B, 1) and res.xs.shape == (K1, B, d)
    assert torch.isfinite(res.samples).all() and abs(res.weights.sum().item() - 1.0) < 1e-4
    for k in ("eval/elbo", "eval/lv_loss", "eval/sample_time"):
        assert math.isfinite(res.metrics[k]), k
    assert math.isfinite(res.log_norm_const_preds["log_norm_const_is"])
    if model.eubo_available and hasattr(model.loss, "compute_eubo"):  # TimeReversalLoss has none (hacking.py:84)
        for k in ("eval/eubo", "eval/log_norm_const_is_f", "eval/effective_sample_size_f"):
            assert math.isfinite(res.metrics[k]), k
        assert 1.0 <= res.metrics["eval/effective_sample_size_f"] <= B + 1e-3
    if solver_type == "dds_orig":
        # cosine grid, dt 0.05, end 6.4: ceil(6.4 / 0.05) = 129 steps in floating point, exactly what the reference's
        # get_timesteps returns; n_steps is ignored for DDS (benchmark_utils.py:184)
        assert K1 == 130
    # the estimators agree with the oracle's formulas on the very same log-weights
    x, rnd, _ = model.loss.simulate(model.eval_ts, model.prior.sample((B,)), model.clipped_target_unnorm_log_prob,
                                    *( [model.reference_distr.log_prob] if hasattr(model, "reference_distr") else []),
                                    **({"initial_log_prob": model.prior.log_prob, "train": False} if solver_type in ("cmcd", "dis_orig") else {}),
                                    seed=5)
    got = type(model.loss).compute_results(rnd, compute_weights=True)
    want = O.compute_results(rnd.cpu())
    assert abs(got.metrics["eval/elbo"] - want["eval/elbo"]) < 1e-4 * max(1, abs(want["eval/elbo"]))
    assert abs(got.log_norm_const_preds["log_norm_const_is"] - want["log_norm_const_is"]) < 1e-3
    assert abs(got.metrics["eval/lv_loss"] - want["eval/lv_loss"]) < 1e-3 * max(1, want["eval/lv_loss"])


@pytest.mark.parametrize("solver_type, model_type", [
    ("vp-ref", "target_informed_zero_init"),
    ("dis_orig", "target_informed_lerp_tempering"),
    ("cmcd", "target_informed_zero_init"),
])
def test_use_ema_with_target_informed_controls(solver_type, model_type, device):
    """make_model(use_ema=True): the EMA AveragedModel deep-copies the control (and, through its bound score methods,
    the target / prior); the copy must keep pointing at the solver's own objects, and objects that have already been
    packed for the kernels (ctypes blocks in their caches) must survive ``copy.deepcopy``."""
    import copy
    from sde_sampler_lrds_b200 import benchmark_utils as BU
    from sde_sampler_lrds_b200 import pack
    d, M = 6, 5
    details = {"sigma": 1.0, **_gmm_ref(d, M), "mean_ref": torch.zeros(d), "var_ref": torch.full((d,), 2.0),
               "mean": torch.zeros(d), "var": torch.full((d,), 4.0)}
    model = BU.make_model(solver_type=solver_type, ref_type="gmm" if solver_type == "vp-ref" else "gaussian" if solver_type == "cmcd" else "default",
                          loss_type="lv", integrator_type="ei" if solver_type == "vp-ref" else "em", model_type=model_type,
                          time_type="uniform", solver_details=details,
                          target_details=BU.make_target_details("many_modes", dim=d, n_modes=M), training_details=TRAIN,
                          n_steps=16, device=str(device), use_ema=True)
    assert isinstance(model.generative_ctrl_ema, torch.optim.swa_utils.AveragedModel)
    assert pack.resolve_ctrl(model.generative_ctrl_ema).target is model.target
    _randomise_last_layers(model)
    model.generative_ctrl_ema.update_parameters(model.generative_ctrl)
    res_ema = model.compute_results(use_ema=True)     # evaluates the EMA copy through the kernels (packs target + weights)
    res_raw = model.compute_results(use_ema=False)
    for r in (res_ema, res_raw):
        assert torch.isfinite(r.samples).all() and math.isfinite(r.metrics["eval/elbo"])
    # a second EMA update + evaluation after the caches are populated, and a deepcopy of a packed control / target
    model.generative_ctrl_ema.update_parameters(model.generative_ctrl)
    assert math.isfinite(model.compute_results(use_ema=True).metrics["eval/elbo"])
    clone = copy.deepcopy(model.generative_ctrl)
    assert clone.base_model._packed == {} and copy.deepcopy(model.target)._lrds_cache == {}
    if solver_type == "cmcd":  # update_prior rebuilds the models (deep copies inside) after everything has been packed
        model.update_prior(torch.zeros(d, device=device), torch.full((d,), 3.0, device=device))
        assert math.isfinite(model.compute_results(use_ema=True).metrics["eval/elbo"])


def test_plan_cache_follows_a_reference_change(device):
    """RDS.change_reference_type swaps the reference objects of a long-lived loss: the next rollout must be packed
    from the NEW reference (never a stale plan), and an in-place edit of the reference's parameters must be seen."""
    from sde_sampler_lrds_b200 import benchmark_utils as BU
    d, M = 6, 5
    details = {"sigma": 1.0, **_gmm_ref(d, M), "mean_ref": torch.zeros(d), "var_ref": torch.full((d,), 2.0)}
    model = BU.make_model(solver_type="vp-ref", ref_type="gmm", loss_type="lv", integrator_type="ei",
                          model_type="target_informed_zero_init", time_type="uniform", solver_details=details,
                          target_details=BU.make_target_details("many_modes", dim=d, n_modes=M), training_details=TRAIN,
                          n_steps=16, device=str(device))
    _randomise_last_layers(model)
    model.compute_results()  # sets eval_ts
    x0 = model.prior.sample((256,))

    def run():
        return model.loss.simulate(model.eval_ts, x0, model.clipped_target_unnorm_log_prob, model.reference_distr.log_prob, seed=3)[1]
    a = run()
    assert torch.equal(a, run())  # cached plan, same result
    model.change_reference_type(ref_type="gaussian", mean=torch.ones(d), var=torch.full((d,), 1.5))
    b = run()
    assert not torch.allclose(a, b)
    for _ in range(4):  # allocate / free look-alike objects: a recycled address must not resurrect an old plan
        model.change_reference_type(ref_type="gaussian", mean=torch.ones(d), var=torch.full((d,), 1.5))
        assert torch.equal(b, run())
    model.reference_score_t.means.add_(0.5)  # in-place edit of the reference's parameters
    model.reference_distr = model.reference_score_t.distr_at(torch.tensor(0.0), model.device)
    assert not torch.allclose(b, run())


def test_make_model_over_the_bracket_two_modes_target(device):
    from sde_sampler_lrds_b200 import benchmark_utils as BU
    model = BU.make_model(solver_type="vp-ref", ref_type="default", loss_type="lv", integrator_type="ei",
                          model_type="target_informed_zero_init", time_type="snr", solver_details={"sigma": 1.0},
                          target_details=BU.make_target_details("bracket_two_modes", dim=6), training_details=TRAIN,
                          n_steps=24, device=str(device))
    _randomise_last_layers(model)
    res = model.compute_results()
    assert res.samples.shape == (TRAIN["eval_batch_size"], 6) and torch.isfinite(res.weights).all()
    assert math.isfinite(res.metrics["eval/elbo"]) and math.isfinite(res.log_norm_const_preds["log_norm_const_is"])


@pytest.mark.parametrize("name", ["dis_many_modes_lerp", "dis_many_modes_langevin", "dis_logreg_lerp", "dis_many_modes_ito"])
def test_drift_model_forward_matches_the_oracle(name, device):
    """``generative_ctrl(t, x)`` (lrds_ctrl_forward) of ScoreCtrl / CancelDriftCtrl / LerpCtrl against the oracle's
    restatement of models/reparam.py, at three times."""
    from tests.cases import CASES, initial_state
    from tests.product_builders import Built
    case = CASES[name]()
    p = case["problem"]
    built = Built(case, device, "fp32")
    _, target_score = O.make_target(p["target"])
    want_fn = O.make_ctrl(p["ctrl"], target_score)
    x = initial_state(case)
    for t in (0.05, 0.5, 0.93):
        t = torch.tensor(t)
        got = built.ctrl(t.to(device), x.to(device)).cpu()
        want = want_fn(t, x)
        err = ((got - want).abs() / want.abs().clamp(min=1.0)).max().item()
        assert err < 1e-4, (name, float(t), err)


@pytest.mark.parametrize("precision", ["tf32x3", "f16x3"])
def test_results_do_not_depend_on_how_particles_are_sharded(device, precision):
    """Counter-based noise keyed by the global particle index: integrating [0, B) in one launch equals integrating
    [0, B/2) and [B/2, B) as two ranks would (particle_offset), bit for bit."""
    from tests.cases import CASES, initial_state
    from tests.product_builders import Built
    case = CASES["ei_many_modes"]()
    x0 = initial_state(case)
    built = Built(case, device, precision)
    xa, ra, _ = built.simulate(x0, None, seed=77, particle_offset=1000)
    h = x0.shape[0] // 2
    xb0, rb0, _ = built.simulate(x0[:h], None, seed=77, particle_offset=1000)
    xb1, rb1, _ = built.simulate(x0[h:], None, seed=77, particle_offset=1000 + h)
    assert torch.equal(xa, torch.cat([xb0, xb1])) and torch.equal(ra, torch.cat([rb0, rb1]))
    xc, rc, _ = built.simulate(x0, None, seed=78, particle_offset=1000)
    assert not torch.equal(ra, rc)


def test_euler_integrator_matches_its_definition(device):
    from sde_sampler_lrds_b200.eq.integrator import EulerIntegrator
    from sde_sampler_lrds_b200.eq.sdes import VP
    sde = VP(diff_coeff_sq_min=0.1, diff_coeff_sq_max=10.0, scale_diff_coeff=1.0, terminal_t=1.0).to(device)
    ts = torch.linspace(0, 1, 17, device=device)
    g = torch.Generator().manual_seed(2)
    x0 = torch.randn(300, 7, generator=g).to(device)
    incs = torch.randn(16, 300, 7, generator=g).to(device)
    k = {"i": 0}

    def bm(s, t):
        k["i"] += 1
        return incs[k["i"] - 1] * torch.sqrt(t - s)
    xs = EulerIntegrator().integrate(sde=sde, ts=ts, x_init=x0, timesteps=ts, bm=bm)
    assert xs.shape == (17, 300, 7) and torch.equal(xs[0], x0)
    h = sde.host()
    x = x0.double().cpu()
    for i, (s, t) in enumerate(zip(ts[:-1].cpu(), ts[1:].cpu())):
        x = x + (h.drift_coeff_t(s).double() * x) * (t - s).double() + h.diff(s).double() * (incs[i].double().cpu() * torch.sqrt(t - s).double())
    assert (xs[-1].double().cpu() - x).abs().max() < 1e-4
    # without an injected Brownian motion: finite, right scale
    xs2 = EulerIntegrator().integrate(sde=sde, ts=ts, x_init=x0, timesteps=ts)
    assert torch.isfinite(xs2).all() and 0.2 < xs2[-1].std().item() < 5.0
