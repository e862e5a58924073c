"""Generates tests/golden/*.pt by running the UNMODIFIED reference (imported from /root/reference).

TEST INFRASTRUCTURE ONLY; runs in the build container only (the GPU box has no /root/reference).
    python -m oracle.make_golden [case ...]

For every case of tests/cases.py it builds the reference's own objects (ScoreCtrl/ClippedCtrl over
FourierMLP/TimeEmbed, VP/PinnedBM/ScaledBM/ControlledLangevinSDE, GMM/PhiFour/LogisticRegression,
the loss classes of sde_sampler/losses/oc.py), loads the case weights with ``load_state_dict``,
replays the case's Brownian increments through ``torch.randn_like`` (the reference has no noise
hook in ``simulate``: losses/oc.py:277,722,1372, eq/sdes.py:537,553,662) and stores x_T, rnd and
the estimator values.  Fixtures hold OUTPUTS only (a few KB each); inputs are regenerated from the
case seed by tests/cases.py.  Import recipe: SURVEY.md 8c (torchquad / torchsde are imported by the
reference at module top but unused on this path, so they are stubbed; nothing else is patched).
"""
from __future__ import annotations

import math
import os
import sys
import types

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
REFERENCE = "/root/reference"


def import_reference():
    for name, attrs in (("torchquad", ["Boole"]), ("torchsde", ["BaseBrownian", "BrownianInterval"])):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, type(a, (), {}))
            sys.modules[name] = m
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import sde_sampler.distr.delta  # noqa: F401
    import sde_sampler.distr.gauss  # noqa: F401
    import sde_sampler.distr.logistic_regression  # noqa: F401
    import sde_sampler.distr.phi_four  # noqa: F401
    import sde_sampler.eq.sdes  # noqa: F401
    import sde_sampler.losses.oc  # noqa: F401
    import sde_sampler.models.mlp  # noqa: F401
    import sde_sampler.models.reparam  # noqa: F401
    import sde_sampler
    return sde_sampler


class ReplayNoise:
    """Temporarily makes torch.randn_like return the recorded increments, one per call."""

    def __init__(self, noise):
        self.noise, self.i = noise, 0

    def __enter__(self):
        self._orig = torch.randn_like

        def fake(x, *a, **k):
            z = self.noise[self.i]
            assert z.shape == x.shape
            self.i += 1
            return z.clone()
        torch.randn_like = fake
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._orig


def build_reference_target(ss, tgt):
    from sde_sampler.distr.gauss import GMM
    from sde_sampler.distr.logistic_regression import LogisticRegression
    from sde_sampler.distr.phi_four import PhiFour
    if tgt["kind"] == "gmm":
        return GMM(dim=tgt["loc"].shape[1], loc=tgt["loc"].clone(), scale=tgt["scale"].clone(),
                   mixture_weights=tgt["weights"].clone(), n_reference_samples=10)
    if tgt["kind"] == "phi4":
        return PhiFour(a=tgt["a"], b=tgt["b"], dim=tgt["dim"], beta=tgt["beta"])
    if tgt["kind"] == "logreg":
        t = LogisticRegression(dim=61, data_type="sonar", intercept_mean=tgt["intercept_mean"],
                               intercept_scale=tgt["intercept_scale"], weight_scale=tgt["weight_scale"])
        # synthetic-shape data instead of the pickled dataset; priors rebuilt for the new width
        p = tgt["X"].shape[1]
        t.X_train, t.y_train = tgt["X"].clone(), tgt["y"].clone()
        t.dim = p + 1
        t.weights_prior = torch.distributions.Independent(
            torch.distributions.Normal(torch.zeros(p), t.weight_scale * torch.ones(p)), 1)
        return t
    raise ValueError(tgt["kind"])


def build_reference_ctrl(c, d, target_score):
    from sde_sampler.models.mlp import FourierMLP, TimeEmbed
    from sde_sampler.distr.gauss import IsotropicGauss
    from sde_sampler.models.reparam import CancelDriftCtrl, ClippedCtrl, LerpCtrl, ScoreCtrl
    num_hidden = sum(1 for k in c["sd"] if k.startswith("base_model.hidden_layer.") and k.endswith(".weight"))
    base = FourierMLP(dim=d, activation=torch.nn.GELU(), num_layers=num_hidden + 2, channels=64)
    if c["kind"] in ("score", "cancel", "lerp"):
        kw = dict(base_model=base, score_model=TimeEmbed(dim_out=1, activation=torch.nn.GELU(), num_layers=4, channels=64),
                  target_score=target_score, detach_score=False, clip_score=c["clip_score"],
                  clip_model=c["clip_model"], scale_score=c["scale_score"])
        if c["kind"] == "score":
            m = ScoreCtrl(**kw)
        elif c["kind"] == "cancel":  # conf/model/langevin_init.yaml
            m = CancelDriftCtrl(sde=build_reference_sde(c["sde"]), langevin_init=True, **kw)
        else:  # conf/model/lerp.yaml (its 'hard_constraint' key is swallowed by ClippedCtrl's **kwargs)
            prior = IsotropicGauss(dim=d, loc=c["prior"]["loc"], scale=c["prior"]["scale"])
            m = LerpCtrl(sde=build_reference_sde(c["sde"]), prior_score=prior.score, hard_constraint=False, **kw)
    else:
        m = ClippedCtrl(base_model=base, clip_model=c["clip_model"])
    missing, unexpected = m.load_state_dict({k: v.clone() for k, v in c["sd"].items()}, strict=True)
    assert not missing and not unexpected
    return m.eval()


def build_reference_sde(s):
    from sde_sampler.eq.sdes import VP, CosineVP, PinnedBM, ScaledBM
    if s["kind"] == "vpcos":
        return CosineVP(c=s["c"], scale_diff_coeff=s["scale"], terminal_t=s["T"])
    if s["kind"] == "vp":
        return VP(diff_coeff_sq_min=s["beta_min"], diff_coeff_sq_max=s["beta_max"], scale_diff_coeff=s["scale"],
                  terminal_t=s["T"])
    if s["kind"] == "pbm":
        return PinnedBM(diff_coeff=s["diff"], terminal_t=s["T"])
    return ScaledBM(diff_coeff=s["diff"], terminal_t=s["T"])


def run_reference(case, grads=False):
    """Returns dict(x_T, rnd[, xs_last]) from the reference for the case; with ``grads`` the training objective
    ``loss(ts, x, ...)`` (method='lv': losses/oc.py:364-394, 1240-1272, 1399-1431) and its autograd gradient with
    respect to every parameter of the control."""
    from sde_sampler.distr.gauss import Gauss, IsotropicGauss
    from sde_sampler.eq.sdes import ControlledLangevinSDE
    from sde_sampler.losses import oc
    from tests.cases import initial_state, noise_for
    p = case["problem"]
    target = build_reference_target(None, p["target"])
    d = target.dim
    ctrl = build_reference_ctrl(p["ctrl"], d, target.score)
    x0 = initial_state(case)
    noise = noise_for(case)
    ts = p["ts"].clone()
    method = p["method"]
    clip_t = p.get("clip_target")

    def target_logp(x):  # TrainableDiff.clipped_target_unnorm_log_prob, solver/oc.py:80-87
        out = target.unnorm_log_prob(x)
        return out if clip_t is None else out.clip(-clip_t, clip_t)

    kw = dict(generative_ctrl=ctrl, generative_ctrl_ema=ctrl, method="lv", max_rnd=1e8)
    if case.get("traj_per_sample", 1) != 1:
        kw.update(method="lv_traj", traj_per_sample=case["traj_per_sample"])
    def train(loss, *args, **kwargs):
        val, _ = loss(ts, x0, target_logp, *args, **kwargs)
        val.backward()
        return {"loss": val.detach(), "grads": {n: (q.grad.clone() if q.grad is not None else torch.zeros_like(q))
                                                for n, q in ctrl.named_parameters()}}

    with (torch.enable_grad() if grads else torch.no_grad()), ReplayNoise(noise) as rp:
        if method in ("em", "ei", "ddpm"):
            sde = build_reference_sde(p["sde"])
            ref = p["ref"]
            if ref["kind"] == "gauss":  # RDS.change_reference_type('gaussian'), solver/oc.py:552-563
                utils = {"x_init": ref["mean"].clone(), "var_init": ref["var"].clone()}
                ref_ctrl = lambda t, x: sde.marginal_score(t=t, x=x, **utils)  # noqa: E731
                ref_distr = sde.marginal_distr(t=torch.tensor(0.0), **utils)
            elif ref["kind"] == "gmm":  # solver/oc.py:564-576
                utils = {"means_init": ref["means"].clone(), "variances_init": ref["variances"].clone(),
                         "weights_init": ref["weights"].clone()}
                ref_ctrl = lambda t, x: sde.marginal_gmm_score(t=t, x=x, **utils)  # noqa: E731
                ref_distr = sde.marginal_gmm_distr(t=torch.tensor(0.0), **utils)
            else:  # PIS.setup_models, solver/oc.py:363-365
                ref_ctrl = None
                ref_distr = sde.marginal_distr(t=sde.terminal_t, x_init=ref["loc"].clone())
            cls = {"em": oc.EMReferenceSDELoss, "ei": oc.EIReferenceSDELoss, "ddpm": oc.DDPMLikeReferenceSDELoss}[method]
            loss = cls(sde=sde, reference_ctrl=ref_ctrl, **kw)
            if grads:
                out = train(loss, ref_distr.log_prob)
            elif case.get("eubo"):
                rnd = loss.compute_eubo(ts, x0.clone(), target_logp, ref_distr.log_prob)
                out = {"rnd": rnd}
            else:
                x, rnd, xs = loss.simulate(ts, x0, target_logp, ref_distr.log_prob, return_traj=True)
                out = {"x_T": x, "rnd": rnd, "xs_mid": xs[len(xs) // 2]}
        elif method == "dds":
            prior = IsotropicGauss(dim=d, loc=p["ref"]["loc"], scale=p["ref"]["scale"])
            loss = oc.ExponentialIntegratorSDELoss(alpha=p["alpha"], sigma=p["sigma"], sde=None, **kw)
            if grads:
                out = train(loss, prior.log_prob)
            else:
                x, rnd, xs = loss.simulate(ts, x0, target_logp, prior.log_prob,
                                           compute_ito_int=case.get("compute_ito_int", True), return_traj=True)
                out = {"x_T": x, "rnd": rnd, "xs_mid": xs[len(xs) // 2]}
        elif method == "dis_ei":  # DiscreteTimeReversalLossEI, oc.py:897-1103 (no config instantiates it; same objects as DIS)
            sde = build_reference_sde(p["sde"])
            prior = IsotropicGauss(dim=d, loc=p["ref"]["loc"], scale=p["ref"]["scale"])
            loss = oc.DiscreteTimeReversalLossEI(sde=sde, **kw)
            if grads:
                out = train(loss, prior.log_prob)
            elif case.get("eubo"):
                out = {"rnd": loss.compute_eubo(ts, x0.clone(), target_logp, prior.log_prob)}
            else:
                x, rnd, xs = loss.simulate(ts, x0, target_logp, initial_log_prob=prior.log_prob, train=False, return_traj=True)
                out = {"x_T": x, "rnd": rnd, "xs_mid": xs[len(xs) // 2]}
        elif method == "dis":  # solver/oc.py:185-262 (Bridge, inference_ctrl=None), eval = TimeReversalLoss.eval, oc.py:1274-1307
            sde = build_reference_sde(p["sde"])
            prior = IsotropicGauss(dim=d, loc=p["ref"]["loc"], scale=p["ref"]["scale"])
            loss = oc.TimeReversalLoss(sde=sde, inference_ctrl=None, **kw)
            if grads:
                out = train(loss, prior.log_prob)
            else:
                x, rnd, xs = loss.simulate(ts, x0, target_logp, initial_log_prob=prior.log_prob, train=False,
                                           compute_ito_int=case.get("compute_ito_int", True), return_traj=True)
                out = {"x_T": x, "rnd": rnd, "xs_mid": xs[len(xs) // 2]}
        elif method == "cmcd":
            pr = p["prior"]
            if pr.get("isotropic"):
                prior = IsotropicGauss(dim=d, loc=float(pr["loc"][0]), scale=float(pr["scale"][0]))
            else:
                prior = Gauss(dim=d, loc=pr["loc"].clone(), scale=pr["scale"].clone())
            sde = ControlledLangevinSDE(target_score=target.score, prior_score=prior.score, diff_coeff=p["diff"],
                                        terminal_t=p["T"], clip_score=p["clip_score"])
            kw2 = dict(kw)
            kw2["max_rnd"] = None
            loss = oc.ControlledLangevinSDELoss(sde=sde, **kw2)
            if grads:
                out = train(loss, prior.log_prob)
            elif case.get("eubo"):
                rnd = loss.compute_eubo(ts, x0.clone(), target_logp, prior.log_prob)
                out = {"rnd": rnd}
            else:
                x, rnd, xs = loss.simulate(ts, x0, target_logp, prior.log_prob, train=False, return_traj=True)
                out = {"x_T": x, "rnd": rnd, "xs_mid": xs[len(xs) // 2]}
        else:
            raise ValueError(method)
        assert rp.i == len(noise), (rp.i, len(noise))
    if grads:
        return out
    rnd = out["rnd"]
    if case.get("eubo"):
        from sde_sampler.additions.hacking import evaluate_eubo  # noqa: F401  (formulas restated below)
        neg = -rnd
        w = torch.softmax(neg, dim=0)
        out["metrics"] = {"eval/log_norm_const_is_f": -rnd.logsumexp(dim=0).item() + math.log(len(w)),
                          "eval/eubo": neg.mean().item(),
                          "eval/effective_sample_size_f": (1.0 / (w ** 2).sum()).item()}
    else:
        res = oc.BaseOCLoss.compute_results(rnd, compute_weights=True)
        m = dict(res.metrics)
        m.update(res.log_norm_const_preds)
        w = res.weights
        m["eval/effective_sample_size"] = (w.sum() ** 2 / (w ** 2).sum()).item()  # eval/metrics.py:134-140
        out["metrics"] = m
    return {k: (v.detach().clone() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}


def run_reference_mala(case):
    """The loop of mcmc_sample(mcmc_type='mala') (experiments/benchmark_utils.py:268-333; that module imports hydra, which
    this image lacks, so its driver loop is replayed here) around the reference's OWN mala_step and heuristics_step_size
    (sde_sampler/additions/mcmc.py), with torch.randn / torch.rand_like replaced by the case's recorded draws."""
    from sde_sampler.additions.mcmc import heuristics_step_size, mala_step
    from tests.cases import mala_inputs
    target = build_reference_target(None, case["target"])
    y_init, noise, unif = mala_inputs(case)

    def target_log_prob_and_grad(y):  # benchmark_utils.py:274-277
        y_ = torch.autograd.Variable(y.clone(), requires_grad=True)
        log_prob_y = target.unnorm_log_prob(y_)
        return log_prob_y.flatten(), torch.autograd.grad(log_prob_y.sum(), y_)[0].detach()
    step_size = case["step_size"] * torch.ones((y_init.shape[0], 1))
    y = torch.autograd.Variable(y_init.clone(), requires_grad=True)
    target_log_prob_y, target_grad_y = target_log_prob_and_grad(y)
    ys, accs = [], []
    k = {"i": 0}
    orig_randn, orig_rand_like = torch.randn, torch.rand_like
    torch.randn = lambda shape, **kw: noise[k["i"]].clone()
    torch.rand_like = lambda t, **kw: unif[k["i"]].clone()
    try:
        for step_id in range(case["n_warmup"] + case["n_steps"]):
            k["i"] = step_id
            y, target_log_prob_y, target_grad_y, log_acc = mala_step(
                y=y, target_log_prob_y=target_log_prob_y, target_grad_y=target_grad_y,
                target_log_prob_and_grad=target_log_prob_and_grad, step_size=step_size)
            step_size = heuristics_step_size(step_size, log_acc)
            accs.append(log_acc.detach().clone())
            if step_id >= case["n_warmup"]:
                ys.append(y.detach().clone())
    finally:
        torch.randn, torch.rand_like = orig_randn, orig_rand_like
    return {"ys": torch.stack(ys), "step_size": step_size.detach().clone(), "log_acc": torch.stack(accs)}


def run_reference_rwmh(case):
    """The same driver loop with mcmc_type='rwmh' (benchmark_utils.py:308-314) around the reference's OWN rwmh_step, whose
    torch.randn_like / torch.rand draws are replaced by the case's recorded ones."""
    from sde_sampler.additions.mcmc import heuristics_step_size, rwmh_step
    from tests.cases import mala_inputs
    target = build_reference_target(None, case["target"])
    y_init, noise, unif = mala_inputs(case)
    step_size = case["step_size"] * torch.ones((y_init.shape[0], 1))
    y = torch.autograd.Variable(y_init.clone(), requires_grad=True)
    target_log_prob_y = target.unnorm_log_prob(y).flatten()
    ys, accs = [], []
    k = {"i": 0}
    orig_randn_like, orig_rand = torch.randn_like, torch.rand
    torch.randn_like = lambda t, **kw: noise[k["i"]].clone()
    torch.rand = lambda shape, **kw: unif[k["i"]].clone()
    try:
        for step_id in range(case["n_warmup"] + case["n_steps"]):
            k["i"] = step_id
            y, target_log_prob_y, log_acc = rwmh_step(y=y, target_log_prob_y=target_log_prob_y,
                                                      target_log_prob=target.unnorm_log_prob, step_size=step_size)
            step_size = heuristics_step_size(step_size, log_acc)
            accs.append(log_acc.detach().clone())
            if step_id >= case["n_warmup"]:
                ys.append(y.detach().clone())
    finally:
        torch.randn_like, torch.rand = orig_randn_like, orig_rand
    return {"ys": torch.stack(ys), "step_size": step_size.detach().clone(), "log_acc": torch.stack(accs)}


def run_reference_aux():
    """Small public interfaces next to the rollout: ``get_timesteps`` grids (utils/common.py:30-82) for every grid kind
    and SDE family the solvers build, and ``EulerIntegrator.integrate`` (eq/integrator.py:84-129) over a VP SDE with
    recorded Brownian increments and off-grid output times."""
    from sde_sampler.utils.common import get_timesteps
    from sde_sampler.eq.integrator import EulerIntegrator
    from tests.cases import AUX_SDES, AUX_GRIDS, euler_case
    out = {"grids": {}}
    for gname, (sde_name, kw) in AUX_GRIDS.items():
        sde = build_reference_sde(AUX_SDES[sde_name]) if sde_name else None
        out["grids"][gname] = get_timesteps(sde=sde, **kw).clone()
    ec = euler_case()
    sde = build_reference_sde(ec["sde"])
    it = iter(ec["noise"])

    def bm(s, t):
        return next(it) * torch.sqrt(t - s)
    integ = EulerIntegrator(dt=ec["dt"], rescale_t=None)
    out["euler_xs"] = integ.integrate(sde, ec["ts"], ec["x0"], bm=bm).clone()
    out["euler_xs_snr"] = None
    it = iter(ec["noise"])
    integ = EulerIntegrator(dt=None, steps=ec["snr_steps"])
    out["euler_xs_snr"] = integ.integrate(sde, ec["ts_snr"], ec["x0"], bm=bm, snr_adapted=True).clone()
    return out


def main(argv):
    from tests.cases import CASES
    import_reference()
    os.makedirs(os.path.join(REPO, "tests", "golden"), exist_ok=True)
    if argv and argv[0] == "--aux":  # python -m oracle.make_golden --aux
        out = run_reference_aux()
        out["torch_version"] = str(torch.__version__)
        path = os.path.join(REPO, "tests", "golden", "aux_grids_euler.pt")
        torch.save(out, path)
        print(f"aux: {len(out['grids'])} grids, euler {tuple(out['euler_xs'].shape)} / snr {tuple(out['euler_xs_snr'].shape)} | "
              f"{os.path.getsize(path)} B")
        return
    if argv and argv[0] == "--mala":  # python -m oracle.make_golden --mala [case ...]
        from tests.cases import MALA_CASES
        for name in argv[1:] or list(MALA_CASES):
            case = MALA_CASES[name]()
            out = run_reference_rwmh(case) if case.get("mcmc_type") == "rwmh" else run_reference_mala(case)
            out["torch_version"] = str(torch.__version__)
            path = os.path.join(REPO, "tests", "golden", name + ".pt")
            torch.save(out, path)
            acc = torch.exp(out["log_acc"].clamp(max=0)).mean().item()
            print(f"{name:24s} ys {tuple(out['ys'].shape)} mean acceptance {acc:.3f} step {out['step_size'].min():.3g}.."
                  f"{out['step_size'].max():.3g} | {os.path.getsize(path)} B")
        return
    if argv and argv[0] == "--grads":  # python -m oracle.make_golden --grads [case ...]
        from tests.cases import GRAD_CASES, grad_case
        for name in argv[1:] or GRAD_CASES:
            out = run_reference(grad_case(name), grads=True)
            out["torch_version"] = str(torch.__version__)
            path = os.path.join(REPO, "tests", "golden", "grad_" + name + ".pt")
            torch.save(out, path)
            gmax = max(float(g.abs().max()) for g in out["grads"].values())
            print(f"grad_{name:24s} loss {float(out['loss']):.6g} max |grad| {gmax:.4g} | {os.path.getsize(path)} B")
        return
    names = argv or list(CASES)
    for name in names:
        case = CASES[name]()
        out = run_reference(case)
        out["torch_version"] = str(torch.__version__)
        path = os.path.join(REPO, "tests", "golden", name + ".pt")
        torch.save(out, path)
        r = out["rnd"]
        print(f"{name:24s} rnd mean {r.mean():+.4e} std {r.std():.3e} finite {bool(torch.isfinite(r).all())} "
              f"| {', '.join(f'{k}={v:.5g}' for k, v in out['metrics'].items())} | {os.path.getsize(path)} B")


if __name__ == "__main__":
    main(sys.argv[1:])
