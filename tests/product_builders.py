"""Builds the PRODUCT objects (sde_sampler_lrds_b200, reference-style classes) for a tests/cases.py problem and
runs the fused CUDA rollout through the same public methods a reference user calls."""
from __future__ import annotations

import torch

from sde_sampler_lrds_b200.distr.gauss import GMM, Gauss, IsotropicGauss
from sde_sampler_lrds_b200.distr.logistic_regression import LogisticRegression
from sde_sampler_lrds_b200.distr.phi_four import PhiFour
from sde_sampler_lrds_b200.eq.sdes import VP, ControlledLangevinSDE, CosineVP, MarginalReference, PinnedBM, ScaledBM
from sde_sampler_lrds_b200.losses import oc
from sde_sampler_lrds_b200.models.mlp import FourierMLP, TimeEmbed
from sde_sampler_lrds_b200.models.reparam import CancelDriftCtrl, ClippedCtrl, LerpCtrl, ScoreCtrl


def build_target(tgt, device):
    if tgt["kind"] == "gmm":
        t = GMM(dim=tgt["loc"].shape[1], loc=tgt["loc"].clone(), scale=tgt["scale"].clone(),
                mixture_weights=tgt["weights"].clone(), n_reference_samples=10)
    elif tgt["kind"] == "phi4":
        t = PhiFour(a=tgt["a"], b=tgt["b"], dim=tgt["dim"], beta=tgt["beta"])
    elif tgt["kind"] == "logreg":
        t = LogisticRegression(X_train=tgt["X"].clone(), y_train=tgt["y"].clone(), intercept_mean=tgt["intercept_mean"],
                               intercept_scale=tgt["intercept_scale"], weight_scale=tgt["weight_scale"])
    else:
        raise ValueError(tgt["kind"])
    return t.to(device)


def build_ctrl(c, d, target, device, prior=None):
    num_hidden = sum(1 for k in c["sd"] if k.startswith("base_model.hidden_layer.") and k.endswith(".weight"))
    base = FourierMLP(dim=d, activation=torch.nn.GELU(), num_layers=num_hidden + 2, channels=64)
    if c["kind"] in ("score", "cancel", "lerp"):
        kw = dict(base_model=base, score_model=TimeEmbed(dim_out=1, activation=torch.nn.GELU(), num_layers=4, channels=64),
                  target_score=target.score, detach_score=False, clip_score=c["clip_score"],
                  clip_model=c["clip_model"], scale_score=c["scale_score"])
        if c["kind"] == "score":
            m = ScoreCtrl(**kw)
        elif c["kind"] == "cancel":
            m = CancelDriftCtrl(sde=build_sde(c["sde"], device), langevin_init=True, **kw)
        else:
            m = LerpCtrl(sde=build_sde(c["sde"], device), prior_score=prior.score, hard_constraint=False, **kw)
    else:
        m = ClippedCtrl(base_model=base, clip_model=c["clip_model"])
    m.load_state_dict({k: v.clone() for k, v in c["sd"].items()}, strict=True)
    return m.to(device).eval()


def build_sde(s, device):
    if s["kind"] == "vpcos":
        sde = CosineVP(c=s["c"], scale_diff_coeff=s["scale"], terminal_t=s["T"])
    elif s["kind"] == "vp":
        sde = VP(diff_coeff_sq_min=s["beta_min"], diff_coeff_sq_max=s["beta_max"], scale_diff_coeff=s["scale"],
                 terminal_t=s["T"])
    elif s["kind"] == "pbm":
        sde = PinnedBM(diff_coeff=s["diff"], terminal_t=s["T"])
    else:
        sde = ScaledBM(diff_coeff=s["diff"], terminal_t=s["T"])
    return sde.to(device)


class Built:
    """loss object + the callables its simulate / compute_eubo take."""

    def __init__(self, case, device, precision=None):
        p = case["problem"]
        self.case, self.device = case, device
        self.target = build_target(p["target"], device)
        d = self.target.dim
        if p["method"] in ("dis", "dis_ei"):  # LerpCtrl interpolates the score of the very prior the rollout starts from
            self.prior = IsotropicGauss(dim=d, loc=p["ref"]["loc"], scale=p["ref"]["scale"]).to(device)
        self.ctrl = build_ctrl(p["ctrl"], d, self.target, device, prior=getattr(self, "prior", None))
        self.ts = p["ts"].clone().to(device)
        method = p["method"]
        kw = dict(generative_ctrl=self.ctrl, generative_ctrl_ema=self.ctrl, method="lv", max_rnd=1e8, precision=precision)
        if case.get("traj_per_sample", 1) != 1:
            kw.update(method="lv_traj", traj_per_sample=case["traj_per_sample"])
        if method in ("em", "ei", "ddpm"):
            self.sde = build_sde(p["sde"], device)
            ref = p["ref"]
            if ref["kind"] == "gauss":
                self.ref = MarginalReference(self.sde, "gaussian", x_init=ref["mean"].clone(), var_init=ref["var"].clone())
                self.ref_distr = self.sde.host().marginal_distr(torch.tensor(0.0), ref["mean"].clone(), ref["var"].clone()).to(device)
            elif ref["kind"] == "gmm":
                self.ref = MarginalReference(self.sde, "gmm", means_init=ref["means"].clone(),
                                             variances_init=ref["variances"].clone(), weights_init=ref["weights"].clone())
                self.ref_distr = self.sde.host().marginal_gmm_distr(torch.tensor(0.0), ref["means"].clone(),
                                                                    ref["variances"].clone(), ref["weights"].clone()).to(device)
            else:
                self.ref = None
                h = self.sde.host()
                self.ref_distr = h.marginal_distr(h.terminal_t, ref["loc"].clone()).to(device)
            cls = {"em": oc.EMReferenceSDELoss, "ei": oc.EIReferenceSDELoss, "ddpm": oc.DDPMLikeReferenceSDELoss}[method]
            self.loss = cls(sde=self.sde, reference_ctrl=self.ref, **kw)
            self.args = (self.target.unnorm_log_prob, self.ref_distr.log_prob)
            self.kwargs = {}
        elif method == "dds":
            self.prior = IsotropicGauss(dim=d, loc=p["ref"]["loc"], scale=p["ref"]["scale"]).to(device)
            self.loss = oc.ExponentialIntegratorSDELoss(alpha=p["alpha"], sigma=p["sigma"], sde=None, **kw)
            self.args = (self.target.unnorm_log_prob, self.prior.log_prob)
            self.kwargs = {"compute_ito_int": case.get("compute_ito_int", True)}
        elif method == "dis":
            self.sde = build_sde(p["sde"], device)
            self.loss = oc.TimeReversalLoss(sde=self.sde, inference_ctrl=None, **kw)
            self.args = (self.target.unnorm_log_prob,)
            self.kwargs = {"initial_log_prob": self.prior.log_prob, "train": False,
                           "compute_ito_int": case.get("compute_ito_int", True)}
        elif method == "dis_ei":
            self.sde = build_sde(p["sde"], device)
            self.loss = oc.DiscreteTimeReversalLossEI(sde=self.sde, **kw)
            self.args = (self.target.unnorm_log_prob, self.prior.log_prob)
            self.kwargs = {"train": False}
        elif method == "cmcd":
            pr = p["prior"]
            if pr.get("isotropic"):
                self.prior = IsotropicGauss(dim=d, loc=float(pr["loc"][0]), scale=float(pr["scale"][0])).to(device)
            else:
                self.prior = Gauss(dim=d, loc=pr["loc"].clone(), scale=pr["scale"].clone()).to(device)
            self.sde = ControlledLangevinSDE(target_score=self.target.score, prior_score=self.prior.score,
                                             diff_coeff=p["diff"], terminal_t=p["T"], clip_score=p["clip_score"]).to(device)
            kw["max_rnd"] = None
            self.loss = oc.ControlledLangevinSDELoss(sde=self.sde, **kw)
            self.args = (self.target.unnorm_log_prob, self.prior.log_prob)
            self.kwargs = {"train": False}
        else:
            raise ValueError(method)

    def simulate(self, x0, noise=None, return_traj=False, **extra):
        return self.loss.simulate(self.ts, x0.to(self.device), *self.args, return_traj=return_traj,
                                  noise=None if noise is None else noise.to(self.device), **self.kwargs, **extra)

    def train_loss(self, x0, noise=None, **extra):
        """[TRAINING] ``loss(ts, x, ...)`` -> (loss, metrics); loss.backward() fills the control's parameter gradients."""
        kw = {"initial_log_prob": self.prior.log_prob} if self.case["problem"]["method"] == "dis" else {}
        return self.loss(self.ts, x0.to(self.device), *self.args, noise=None if noise is None else noise.to(self.device),
                         **kw, **extra)

    def compute_eubo(self, x0, noise=None, **extra):
        return self.loss.compute_eubo(self.ts, x0.to(self.device).clone(), *self.args,
                                      noise=None if noise is None else noise.to(self.device), **extra)
