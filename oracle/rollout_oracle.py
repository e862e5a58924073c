"""CPU oracle for the batched trajectory rollout of sde_sampler_lrds.

TEST INFRASTRUCTURE ONLY.  This module is a plain torch-on-CPU restatement of the
reference's integrate-and-weight loop.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the
product package ``sde_sampler_lrds_b200`` never does (it fails loudly without its
CUDA library).

Parity pin: the reference ships no golden vectors for this path (its only test file
``tests/distr_eval.py`` is stale), so the oracle is pinned
against OUTPUTS OF THE REFERENCE ITSELF, produced in the build container by
``oracle/make_golden.py`` (imports /root/reference with two import stubs) and
committed under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them.

Every function cites the reference file:line (relative to /root/reference/) whose
arithmetic it restates.  All time-only quantities are 0-dim tensors, like the
reference's ``for s, t in zip(ts[:-1], ts[1:])`` loop variables.  The Brownian
increments are an explicit ``noise[K, B, d]`` tensor of standard normals (the
reference draws them inline with ``torch.randn_like`` - one draw per step).
"""
from __future__ import annotations

import math
from typing import Callable

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# drift networks: sde_sampler/models/mlp.py, sde_sampler/models/reparam.py
# --------------------------------------------------------------------------------------


def _linear(sd: dict, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def time_embed(sd: dict, prefix: str, t: torch.Tensor) -> torch.Tensor:
    """TimeEmbed.forward, sde_sampler/models/mlp.py:85-96 (ctor 57-83).

    ``timestep_coeff`` is a non-persistent buffer linspace(0.1, 100, channels) (mlp.py:70-73).
    Activation is exact-erf GELU (conf/model/base/time_embed.yaml).  Returns (rows(t), dim_out).
    """
    phase = sd[prefix + "timestep_phase"]
    channels = phase.shape[1]
    coeff = torch.linspace(start=0.1, end=100, steps=channels).unsqueeze(0).to(phase.dtype)
    t = t.reshape(-1, 1).to(phase.dtype)
    emb = torch.cat([torch.sin(coeff * t + phase), torch.cos(coeff * t + phase)], dim=1)
    i = 0
    while f"{prefix}hidden_layer.{i}.weight" in sd:
        emb = F.gelu(_linear(sd, f"{prefix}hidden_layer.{i}", emb))
        i += 1
    return _linear(sd, prefix + "out_layer", emb)


def fourier_mlp(sd: dict, prefix: str, t: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """FourierMLP.forward, sde_sampler/models/mlp.py:135-143 (ctor 99-133)."""
    embed_t = time_embed(sd, prefix + "timestep_embed.", t).expand(x.shape[0], -1)
    embed = _linear(sd, prefix + "input_embed", x) + embed_t
    i = 0
    while f"{prefix}hidden_layer.{i}.weight" in sd:
        embed = _linear(sd, f"{prefix}hidden_layer.{i}", F.gelu(embed))
        i += 1
    return _linear(sd, prefix + "out_layer", F.gelu(embed))


def clip(x: torch.Tensor, max_norm) -> torch.Tensor:
    """clip_and_log, sde_sampler/utils/common.py:85-112 (the logging branch is commented out)."""
    return x if max_norm is None else x.clip(min=-1.0 * max_norm, max=max_norm)


def make_ctrl(ctrl: dict, target_score: Callable | None, dtype=torch.float32) -> Callable:
    """ClippedCtrl.forward (models/reparam.py:33-43) / ScoreCtrl.forward (91-117) / CancelDriftCtrl.forward (131-147,
    use_rescaling=True) / LerpCtrl.forward (170-199, hard_constrain=False).

    ctrl = {"kind": "clipped"|"score"|"cancel"|"lerp", "sd": state_dict, "clip_model", "clip_score", "scale_score"
            [, "sde": sde dict (cancel, lerp)] [, "prior": {"loc", "scale"} of the IsotropicGauss prior (lerp)]}.
    """
    sd = ctrl["sd"]
    kind = ctrl["kind"]
    sde = make_sde(ctrl["sde"], dtype) if kind in ("cancel", "lerp") else None

    def forward(t: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        base = clip(fourier_mlp(sd, "base_model.", t, x), ctrl.get("clip_model"))
        if kind == "clipped":
            return base
        raw = target_score(x)
        if kind == "lerp":  # clipped_interpolated_score, reparam.py:170-183; IsotropicGauss.score, distr/gauss.py:764-766
            pr = ctrl["prior"]
            raw = torch.lerp((pr["loc"] - x) / pr["scale"] ** 2, raw, t / sde.T)
        score = ctrl.get("scale_score", 1.0) * clip(raw, ctrl.get("clip_score"))
        if any(k.startswith("score_model.") for k in sd):
            score = score * clip(time_embed(sd, "score_model.", t), ctrl.get("clip_model"))
        if kind == "cancel":  # reparam.py:135-145
            return base + (sde.drift_coeff(t) * x) / sde.diff(t) + 0.5 * sde.diff(t) * score
        if kind == "lerp":  # reparam.py:199
            return base + sde.diff(t) * score
        return base + score

    return forward


# --------------------------------------------------------------------------------------
# distributions: sde_sampler/distr/{gauss,phi_four,logistic_regression}.py
# --------------------------------------------------------------------------------------


def log_prob_gaussian(x, mean, variance):
    """log_prob_gaussian, sde_sampler/distr/gauss.py:67-73 -> (B, M)."""
    lp = -0.5 * torch.sum(torch.square(x.unsqueeze(1) - mean.unsqueeze(0)) / variance.unsqueeze(0), dim=-1)
    lp = lp - 0.5 * mean.shape[-1] * math.log(2.0 * math.pi)
    lp = lp - 0.5 * torch.log(variance).sum(dim=-1).unsqueeze(0)
    return lp


def score_mog(x, weights, means, variances):
    """score_mog, sde_sampler/distr/gauss.py:97-107 (weights normalised; the in-place
    normalisation of the caller's buffer at line 100 is idempotent and not reproduced)."""
    weights = weights / weights.sum()
    probs = torch.softmax(torch.log(weights.unsqueeze(0)) + log_prob_gaussian(x, means, variances), dim=-1)
    return -torch.sum(probs.unsqueeze(-1) * (x.unsqueeze(1) - means.unsqueeze(0)) / variances.unsqueeze(0), dim=1)


def score_gauss(x, means, variances):
    """score_gauss, sde_sampler/distr/gauss.py:124-126."""
    return -(x - means) / variances


def gmm_log_prob(x, loc, scale, weights):
    """GMM.unnorm_log_prob, sde_sampler/distr/gauss.py:202-221: MixtureSameFamily(Categorical(w),
    Independent(Normal(loc, scale), 1)).log_prob; a single component without weights is
    Independent(Normal) (line 205-208).  Returns (B, 1)."""
    var = scale ** 2
    comp = (-((x.unsqueeze(1) - loc.unsqueeze(0)) ** 2) / (2 * var.unsqueeze(0))
            - scale.log().unsqueeze(0) - math.log(math.sqrt(2 * math.pi))).sum(-1)
    if weights is None:
        return comp[:, :1]
    logw = torch.log_softmax(torch.log(weights / weights.sum()), dim=-1)
    return torch.logsumexp(comp + logw.unsqueeze(0), dim=-1, keepdim=True)


def isotropic_gauss_log_prob(x, loc: float, scale: float):
    """IsotropicGauss.unnorm_log_prob, sde_sampler/distr/gauss.py:757-762 (log_norm_const = 0)."""
    var = torch.as_tensor(scale, dtype=x.dtype) ** 2
    norm_const = -0.5 * x.shape[-1] * (2.0 * math.pi * var).log()
    return norm_const - 0.5 * torch.sum((x - loc) ** 2, dim=-1, keepdim=True) / var


def phi4_U(x, a: float, b: float):
    """PhiFour.U + V, sde_sampler/distr/phi_four.py:45-79 (dim_phys=1, Dirichlet-0, no tilt)."""
    coef = a * x.shape[-1]
    V = ((1 - x ** 2) ** 2 / 4 + b * x).sum(-1) / coef
    x_ = F.pad(x, (1, 1), mode="constant", value=0.0)
    grad_term = ((x_[:, 1:] - x_[:, :-1]) ** 2 / 2).sum(-1)
    return grad_term * coef + V


def phi4_grad_U(x, a: float, b: float):
    """PhiFour.grad_U, sde_sampler/distr/phi_four.py:81-90."""
    coef = a * x.shape[-1]
    ret = (b - x * (1.0 - torch.square(x))) / coef
    ret[:, 1:-1] += coef * (2.0 * x[:, 1:-1] - x[:, 2:] - x[:, :-2])
    ret[:, 0] += coef * (2.0 * x[:, 0] - x[:, 1])
    ret[:, -1] += coef * (2.0 * x[:, -1] - x[:, -2])
    return ret


LOGREG_EPS = {torch.float32: torch.finfo(torch.float32).eps, torch.float64: torch.finfo(torch.float64).eps}


def logreg_log_prob(params, X, y, weight_scale, intercept_mean, intercept_scale, threshold=1e-8):
    """LogisticRegression.posterior_log_prob, sde_sampler/distr/logistic_regression.py:41-61
    (use_intercept=True).  Returns (B,)."""
    weights, intercept = params[..., :-1], params[..., -1]
    ws = torch.as_tensor(weight_scale, dtype=params.dtype)
    prior = (-(weights ** 2) / (2 * ws ** 2) - ws.log() - math.log(math.sqrt(2 * math.pi))).sum(-1)
    isc = torch.as_tensor(intercept_scale, dtype=params.dtype)
    prior = prior + (-((intercept - intercept_mean) ** 2) / (2 * isc ** 2) - isc.log()
                     - math.log(math.sqrt(2 * math.pi)))
    probs = torch.special.expit(torch.matmul(X, weights.T).T + intercept.unsqueeze(-1))
    probs = torch.clip(probs, threshold, 1.0 - threshold)
    eps = LOGREG_EPS[params.dtype]
    ps = probs.clamp(min=eps, max=1 - eps)  # torch.distributions.utils.probs_to_logits(is_binary=True)
    logits = torch.log(ps) - torch.log1p(-ps)
    ll = -F.binary_cross_entropy_with_logits(logits, y.unsqueeze(0).expand(logits.shape[0], -1),
                                             reduction="none").sum(-1)
    return ll + prior


def logreg_score(params, X, y, weight_scale, intercept_mean, intercept_scale, threshold=1e-8):
    """Score the solvers actually use for LogisticRegression: ``score`` is NOT overridden
    (logistic_regression.py:91-92 commented out) so it is autograd of posterior_log_prob through
    the two clamps (distr/base.py:146-154).  Closed form: prior terms +
    sum_n m_n (y_n - p_n) x_n with m_n = [thr <= sigma(z_n) <= 1-thr] & [eps <= p_n <= 1-eps] (inclusive)."""
    weights, intercept = params[..., :-1], params[..., -1]
    z = torch.matmul(X, weights.T).T + intercept.unsqueeze(-1)
    sig = torch.special.expit(z)
    eps = LOGREG_EPS[params.dtype]
    m1 = (sig >= threshold) & (sig <= 1.0 - threshold)
    p = torch.clip(sig, threshold, 1.0 - threshold)
    m2 = (p >= eps) & (p <= 1 - eps)
    ps = p.clamp(min=eps, max=1 - eps)
    # d/dz [ y log ps + (1-y) log(1-ps) ] through logits = log ps - log1p(-ps), bce_with_logits
    # = (y - sigmoid(logit(ps))) * dlogit/dps * dps/dsig * sig(1-sig); sigmoid(logit(ps)) = ps up to rounding
    g = (y.unsqueeze(0) - ps) * (sig * (1 - sig)) / (ps * (1 - ps)) * (m1 & m2)
    gw = g @ X - weights / weight_scale ** 2
    gi = g.sum(-1, keepdim=True) - (intercept.unsqueeze(-1) - intercept_mean) / intercept_scale ** 2
    return torch.cat([gw, gi], dim=-1)


def make_target(target: dict):
    """Returns (unnorm_log_prob(x)->(B,1), score(x)->(B,d)) for a target description dict."""
    kind = target["kind"]
    if kind == "gmm":
        loc, scale, w = target["loc"], target["scale"], target.get("weights")

        def logp(x):
            return gmm_log_prob(x, loc, scale, w)

        def score(x):
            if w is None:  # Gauss.score, distr/gauss.py:627-629
                return score_gauss(x, loc, torch.square(scale))
            return score_mog(x, w, loc, torch.square(scale))  # GMM.score, distr/gauss.py:241-243
        return logp, score
    if kind == "phi4":
        a, b, beta = target["a"], target["b"], target["beta"]
        # PhiFour.unnorm_log_prob / score, distr/phi_four.py:92-96
        return (lambda x: -beta * phi4_U(x, a, b).unsqueeze(-1)), (lambda x: -beta * phi4_grad_U(x, a, b))
    if kind == "logreg":
        args = (target["X"], target["y"], target["weight_scale"], target["intercept_mean"],
                target["intercept_scale"])
        return (lambda x: logreg_log_prob(x, *args).unsqueeze(-1)), (lambda x: logreg_score(x, *args))
    raise ValueError(kind)


# --------------------------------------------------------------------------------------
# SDE scalar algebra: sde_sampler/eq/sdes.py
# --------------------------------------------------------------------------------------


class VP:
    """VP, sde_sampler/eq/sdes.py:427-555 (all buffers float32, lines 438-454)."""

    def __init__(self, beta_min=0.1, beta_max=20.0, scale=1.0, T=1.0, dtype=torch.float32):
        self.bmin = torch.tensor(beta_min, dtype=dtype)
        self.bmax = torch.tensor(beta_max, dtype=dtype)
        self.c = torch.tensor(scale, dtype=dtype)
        self.T = torch.tensor(T, dtype=dtype)

    def beta(self, t):  # _diff_coeff_sq_t, 456-459
        return torch.lerp(self.bmin, self.bmax, t / self.T)

    def drift_coeff(self, t):  # 461-463
        return -0.5 * self.beta(t)

    def diff(self, t):  # 465-467
        return self.c * torch.sqrt(self.beta(t))

    def int_drift_coeff(self, s, t):  # int_drift_coeff_t, 469-477
        return -0.25 * (self.beta(t) + self.beta(s)) * (t - s)

    def alpha_(self, t):  # 490-493
        return self.bmin * t + (0.5 * t ** 2 / self.T) * (self.bmax - self.bmin)

    def transition_params(self, s, t):  # 495-507
        lam = 1.0 - torch.exp(self.alpha_(s) - self.alpha_(t))
        return torch.sqrt(1.0 - lam), self.c ** 2 * lam

    def s(self, t):  # 509-511
        return torch.exp(-0.5 * self.alpha_(t))

    def sigma_sq(self, t):  # 513-515
        return -self.c ** 2 * (1.0 - (1.0 / self.s(t) ** 2))

    def omega(self, tk, tk1):  # 517-520
        return 4.0 * self.c ** 2 * torch.tanh((self.alpha_(self.T - tk) - self.alpha_(self.T - tk1)) / 4.0)

    def lambda_(self, tk, tk1):  # 522-524
        return torch.exp(self.alpha_(self.T - tk) - self.alpha_(self.T - tk1)) - 1.0

    def omega_ddpm(self, tk, tk1):  # 526-530
        lk = 1.0 - torch.exp(-self.alpha_(self.T - tk))
        lk1 = 1.0 - torch.exp(-self.alpha_(self.T - tk1))
        return self.c ** 2 * (lk / lk1) * self.lambda_(tk, tk1)

    def ei_step(self, x, tk, tk1, s, z):  # ei_integration_step, 532-539
        lam = self.lambda_(tk, tk1)
        ret = torch.sqrt(1.0 + lam) * x + 2.0 * self.c ** 2 * (torch.sqrt(1.0 + lam) - 1.0) * s
        return ret + self.c * torch.sqrt(lam) * z

    def ddpm_step(self, x, tk, tk1, s, z):  # ddpm_integration_step, 541-555
        T = self.T
        lam = self.lambda_(tk, tk1)
        lam__ = 1.0 - torch.exp(self.alpha_(T - tk1) - self.alpha_(T - tk))
        lk = 1.0 - torch.exp(-self.alpha_(T - tk))
        lk1 = 1.0 - torch.exp(-self.alpha_(T - tk1))
        diff_alpha = (self.alpha_(T - tk) - self.alpha_(T - tk1)) / 2.0
        var = self.c ** 2 * lam__ * (lk1 / lk)
        mean = torch.sqrt(1.0 + lam) * x + 2.0 * self.c ** 2 * torch.sinh(diff_alpha) * s
        return mean + torch.sqrt(var) * z


class CosineVP(VP):
    """CosineVP, sde_sampler/eq/sdes.py:558-594: VP with beta(t) and alpha(t) of the cosine schedule."""

    def __init__(self, c=0.008, scale=1.0, T=1.0, dtype=torch.float32):
        super().__init__(0.1, 20.0, scale, T, dtype)
        self.cc = torch.tensor(c, dtype=dtype)

    def beta(self, t):  # _diff_coeff_sq_t, 577-580
        return torch.pi * torch.tan(0.5 * torch.pi * ((t / self.T) + self.cc) / (1.0 + self.cc)) / (self.T * (1.0 + self.cc))

    def alpha_(self, t):  # 592-594
        return -2.0 * torch.log(torch.cos(0.5 * torch.pi * ((t / self.T) + self.cc) / (1.0 + self.cc)))


class PinnedBM:
    """PinnedBM, sde_sampler/eq/sdes.py:597-678."""

    def __init__(self, diff=math.sqrt(0.2), T=5.0, dtype=torch.float32):
        self.sig = torch.tensor(diff, dtype=dtype)
        self.T = torch.tensor(T, dtype=dtype)

    def drift_coeff(self, t):  # 609-611
        return -1.0 / (self.T - t)

    def diff(self, t):  # 613-615
        return self.sig

    def transition_params(self, s, t):  # 628-639
        mf = (self.T - t) / (self.T - s)
        return mf, mf * (t - s) * self.sig ** 2

    def s(self, t):  # 641-643
        return (self.T - t) / self.T

    def sigma_sq(self, t):  # 645-647
        return self.sig ** 2 * self.T * t / (self.T - t)

    def omega(self, tk, tk1):  # 649-651
        return self.sig ** 2 * (tk / tk1) * (tk1 - tk)

    def omega_ddpm(self, tk, tk1):  # 653-656
        return self.sig ** 2 * ((self.T - tk) / (self.T - tk1)) * (tk1 - tk)

    def ei_step(self, x, tk, tk1, s, z):  # 658-666
        ret = (tk1 / tk) * x + self.sig ** 2 * (tk1 - tk) * s
        var = self.sig ** 2 * (tk1 / tk) * (tk1 - tk)
        return ret + torch.sqrt(var) * z

    def ddpm_step(self, x, tk, tk1, s, z):  # 668-678
        var = self.sig ** 2 * ((self.T - tk1) / (self.T - tk)) * (tk1 - tk)
        mean = (tk1 / tk) * x + self.sig ** 2 * (tk1 - tk) * s
        return mean + torch.sqrt(var) * z


class ScaledBM:
    """ScaledBM(ConstOU), sde_sampler/eq/sdes.py:354-424 (drift 0, constant diffusion)."""

    def __init__(self, diff=math.sqrt(0.2), T=5.0, dtype=torch.float32):
        self.sig = torch.tensor(diff, dtype=dtype)
        self.T = torch.tensor(T, dtype=dtype)

    def drift_coeff(self, t):  # 378-380 with drift_coeff = 0
        return -torch.zeros((), dtype=self.sig.dtype)

    def diff(self, t):  # 382-384
        return self.sig

    def s(self, t):  # 418-420
        return torch.ones_like(t)

    def sigma_sq(self, t):  # 422-424
        return self.sig ** 2 * t


def make_sde(sde: dict | None, dtype=torch.float32):
    if sde is None:
        return None
    kind = sde["kind"]
    if kind == "vp":
        return VP(sde["beta_min"], sde["beta_max"], sde.get("scale", 1.0), sde.get("T", 1.0), dtype)
    if kind == "vpcos":
        return CosineVP(sde.get("c", 0.008), sde.get("scale", 1.0), sde.get("T", 1.0), dtype)
    if kind == "pbm":
        return PinnedBM(sde["diff"], sde["T"], dtype)
    if kind == "bm":
        return ScaledBM(sde["diff"], sde["T"], dtype)
    raise ValueError(kind)


def marginal_params(sde, t, x_init, var_init=None):
    """OU.marginal_params, sde_sampler/eq/sdes.py:208-248 (diagonal branch)."""
    loc = sde.s(t) * x_init
    var = sde.s(t) ** 2 * sde.sigma_sq(t)
    if var_init is not None:
        var = var + sde.s(t) ** 2 * var_init
    return loc, var


def make_reference(ref: dict | None, sde):
    """RDS.change_reference_type, sde_sampler/solver/oc.py:513-588 ('default'/'gaussian'/'gmm',
    diagonal) and PIS.setup_models, solver/oc.py:363-365.  Returns
    (reference_ctrl(t, x) | None, reference_log_prob(x) -> (B, 1))."""
    kind = ref["kind"]
    if kind == "gauss":  # sdes.py:250-279 at t: marginal_score; at t=0: marginal_distr -> Gauss.log_prob
        x_init, var_init = ref["mean"], ref["var"]

        def ctrl(t, x):
            loc, var = marginal_params(sde, t, x_init, var_init)
            return score_gauss(x, loc, var)
        loc0, var0 = marginal_params(sde, torch.zeros((), dtype=x_init.dtype), x_init, var_init)
        return ctrl, (lambda x: gmm_log_prob(x, loc0.unsqueeze(0), var0.sqrt().unsqueeze(0), None))
    if kind == "gmm":  # sdes.py:281-345
        means, variances, w = ref["means"], ref["variances"], ref["weights"]

        def ctrl(t, x):
            loc, var = marginal_params(sde, t, means, variances)
            return score_mog(x, w, loc, var)
        loc0, var0 = marginal_params(sde, torch.zeros((), dtype=means.dtype), means, variances)
        return ctrl, (lambda x: gmm_log_prob(x, loc0, torch.sqrt(var0), w))
    if kind == "pis":  # N(prior.loc, sigma^2 T): sde.marginal_distr(T, x_init=prior.loc), no reference_ctrl
        loc, var = marginal_params(sde, sde.T, ref["loc"])
        var = var * torch.ones_like(ref["loc"])
        return None, (lambda x: gmm_log_prob(x, loc.unsqueeze(0), var.sqrt().unsqueeze(0), None))
    if kind == "iso":  # DDS: reference_distr = prior = IsotropicGauss(scale=sigma), solver/oc.py:445-446
        return None, (lambda x: isotropic_gauss_log_prob(x, ref.get("loc", 0.0), ref["scale"]))
    raise ValueError(kind)


# --------------------------------------------------------------------------------------
# rollouts: sde_sampler/losses/oc.py
# --------------------------------------------------------------------------------------


def _running_cost(g, lv):
    """(generative_ctrl, sde_ctrl, cost density): with ``lv`` the training form of the losses with change_sde_ctrl=True
    (BaseOCLoss.generative_and_sde_ctrl, losses/oc.py:83-103, no control noise / dropout: the SDE follows the detached
    control and the cost is g (sde_ctrl - g / 2), e.g. oc.py:272-273, 489-490); else 0.5 g^2 (oc.py:275, 492)."""
    if lv:
        u = g.detach()
        return u, (g * (u - 0.5 * g)).sum(dim=-1, keepdim=True)
    return g, 0.5 * (g ** 2).sum(dim=-1, keepdim=True)


def simulate_em(ts, x, noise, ctrl, sde, reference_ctrl, terminal_unnorm_log_prob, reference_log_prob,
                return_traj=False, lv=False):
    """EMReferenceSDELoss.simulate, sde_sampler/losses/oc.py:218-296 (use_rescaling=True; ``lv`` = change_sde_ctrl)."""
    rnd = 0.0
    T = ts[-1]
    xs = [x] if return_traj else None
    for k, (s, t) in enumerate(zip(ts[:-1], ts[1:])):
        g = ctrl(T - s, x)
        u, cost = _running_cost(g, lv)
        sde_diff = sde.diff(T - s)
        dt = t - s
        rnd = rnd + cost * dt
        db = noise[k] * dt.sqrt()
        drift_ = -(sde.drift_coeff(T - s) * x)
        if reference_ctrl is not None:
            drift_ = drift_ + torch.square(sde_diff) * reference_ctrl(T - s, x)
        x = x + (drift_ + sde_diff * u) * dt + sde_diff * db
        rnd = rnd + (g * db).sum(dim=-1, keepdim=True)
        if return_traj:
            xs.append(x)
    rnd = rnd + reference_log_prob(x).view((-1, 1)) - terminal_unnorm_log_prob(x)
    return x, rnd, (torch.stack(xs) if return_traj else None)


def simulate_ei(ts, x, noise, ctrl, sde, reference_ctrl, terminal_unnorm_log_prob, reference_log_prob,
                return_traj=False, ddpm=False, lv=False):
    """EIReferenceSDELoss.simulate, losses/oc.py:444-510; with ddpm=True
    DDPMLikeReferenceSDELoss.simulate, losses/oc.py:584-651."""
    rnd = 0.0
    T = ts[-1]
    xs = [x] if return_traj else None
    omega = sde.omega_ddpm if ddpm else sde.omega
    step = sde.ddpm_step if ddpm else sde.ei_step
    for k, (s, t) in enumerate(zip(ts[:-1], ts[1:])):
        g = ctrl(T - s, x)
        u, cost = _running_cost(g, lv)
        rnd = rnd + (omega(s, t) * cost if lv else 0.5 * omega(s, t) * (g ** 2).sum(dim=-1, keepdim=True))
        z = noise[k]
        x = step(x, s, t, reference_ctrl(T - s, x) + u, z)
        rnd = rnd + torch.sqrt(omega(s, t)) * (g * z).sum(dim=-1, keepdim=True)
        if return_traj:
            xs.append(x)
    rnd = rnd + reference_log_prob(x).view((-1, 1)) - terminal_unnorm_log_prob(x)
    return x, rnd, (torch.stack(xs) if return_traj else None)


def simulate_dds(ts, x, noise, ctrl, alpha, sigma, terminal_unnorm_log_prob, reference_log_prob,
                 compute_ito_int=True, return_traj=False, lv=False):
    """ExponentialIntegratorSDELoss.simulate, losses/oc.py:1319-1397 (forward time; ``lv`` = change_sde_ctrl)."""
    rnd = 0.0
    xs = [x] if return_traj else None
    for k, (s, t) in enumerate(zip(ts[:-1], ts[1:])):
        g = ctrl(s, x)
        u, running_cost = _running_cost(g, lv)
        dt = t - s
        beta_k = torch.clip(alpha * dt.sqrt(), 0, 1)
        alpha_k = torch.sqrt(1.0 - beta_k ** 2)
        rnd = rnd + beta_k ** 2 * sigma ** 2 * running_cost
        z = noise[k]
        x = x * alpha_k + (beta_k ** 2) * (sigma ** 2) * u + sigma * beta_k * z
        if compute_ito_int:
            rnd = rnd + (sigma * g * z * beta_k).sum(dim=-1, keepdim=True)
        if return_traj:
            xs.append(x)
    rnd = rnd + reference_log_prob(x).view((-1, 1)) - terminal_unnorm_log_prob(x)
    return x, rnd, (torch.stack(xs) if return_traj else None)


def simulate_dis(ts, x, noise, ctrl, sde, terminal_unnorm_log_prob, initial_log_prob, compute_ito_int=True,
                 return_traj=False, lv=False):
    """TimeReversalLoss.simulate, losses/oc.py:1133-1238, as DIS evaluates it (solver/oc.py:185-262 Bridge with
    inference_ctrl=None; eval passes train=False, change_sde_ctrl=False, use_rescaling=True): the control is taken at
    the loop time s (not T - s), the drift is +sde.drift(s, x), the log-weight starts at the prior log-density and
    carries the divergence integral OU.drift_div_int = d * int_drift_coeff_t (eq/sdes.py:137-141)."""
    rnd = initial_log_prob(x)
    d = x.shape[-1]
    xs = [x] if return_traj else None
    for k, (s, t) in enumerate(zip(ts[:-1], ts[1:])):
        g = ctrl(s, x)
        u, cost = _running_cost(g, lv)
        sde_diff = sde.diff(s)
        dt = t - s
        rnd = rnd + cost * dt
        if not lv:  # `if not train` (oc.py:1217-1218): the training rollout leaves the divergence integral out
            rnd = rnd - sde.int_drift_coeff(s, t) * d
        db = noise[k] * dt.sqrt()
        x = x + (sde.drift_coeff(s) * x + sde_diff * u) * dt + sde_diff * db
        if compute_ito_int:
            rnd = rnd + (g * db).sum(dim=-1, keepdim=True)
        if return_traj:
            xs.append(x)
    rnd = rnd - terminal_unnorm_log_prob(x)
    return x, rnd, (torch.stack(xs) if return_traj else None)


def simulate_dis_ei(ts, x, noise, ctrl, sde, terminal_unnorm_log_prob, initial_log_prob, return_traj=False, lv=False):
    """DiscreteTimeReversalLossEI.simulate, losses/oc.py:906-978 (eval: train=False; the lv training form starts at the
    prior log-density as well, 925-929)."""
    rnd = initial_log_prob(x)
    T = ts[-1]
    xs = [x] if return_traj else None
    for k, (s, t) in enumerate(zip(ts[:-1], ts[1:])):
        g = ctrl(T - s, x)
        u, cost = _running_cost(g, lv)
        rnd = rnd + (sde.omega(s, t) * cost if lv else 0.5 * sde.omega(s, t) * (g ** 2).sum(dim=-1, keepdim=True))
        z = noise[k]
        x = sde.ei_step(x, s, t, u, z)
        rnd = rnd + torch.sqrt(sde.omega(s, t)) * (g * z).sum(dim=-1, keepdim=True)
        if return_traj:
            xs.append(x)
    rnd = rnd - terminal_unnorm_log_prob(x)
    return x, rnd, (torch.stack(xs) if return_traj else None)


def eubo_dis_ei(ts, x, noise, ctrl, sde, terminal_unnorm_log_prob, initial_log_prob):
    """DiscreteTimeReversalLossEI.compute_eubo, losses/oc.py:980-1036."""
    x = x.clone()
    rnd = -terminal_unnorm_log_prob(x)
    T = ts[-1]
    times_s, times_t = ts[:-1].flip((0,)), ts[1:].flip((0,))
    mean_factors, var_factors = sde.transition_params(T - times_t, T - times_s)
    std_factors = var_factors.sqrt()
    for i, (s, t) in enumerate(zip(times_s, times_t)):
        z = noise[i]
        x = x * mean_factors[i]
        x = x + std_factors[i] * z
        g = ctrl(T - s, x)
        rnd = rnd - (0.5 * g ** 2).sum(dim=-1, keepdim=True) * sde.omega(s, t)
        rnd = rnd - (g * z).sum(dim=-1, keepdim=True) * torch.sqrt(sde.omega(s, t))
    return rnd + initial_log_prob(x)


def langevin_drift(t, x, target_score, prior_score, diff, T, clip_score):
    """ControlledLangevinSDE.drift, sde_sampler/eq/sdes.py:101-110."""
    drift = target_score(x) * (t / T) + prior_score(x) * (1.0 - t / T)
    drift = drift * (0.5 * diff ** 2)
    return clip(drift, clip_score)


def simulate_cmcd(ts, x, noise, ctrl, drift, diff, terminal_unnorm_log_prob, initial_log_prob,
                  return_traj=False, lv=False):
    """ControlledLangevinSDELoss.simulate, losses/oc.py:666-755 (train=False, use_rescaling=True,
    change_sde_ctrl=False).  (u_t, drift_t) of step k are re-evaluated at step k+1 as (u_s, drift_s)
    exactly like the reference (no caching here: the oracle keeps the reference's evaluation count).  ``lv`` = the
    training form (change_sde_ctrl=True, train=True with method 'lv': the SDE follows the detached control, 705-706, 739)."""
    rnd = initial_log_prob(x)
    xs = [x] if return_traj else None
    for k, (s, t) in enumerate(zip(ts[:-1], ts[1:])):
        u_s = ctrl(s, x)
        sde_ctrl = u_s.detach() if lv else u_s
        dt = t - s
        db = dt.sqrt() * noise[k]
        drift_s = drift(s, x)
        y = x + (drift_s + sde_ctrl * diff) * dt + diff * db
        drift_t = drift(t, y)
        u_t = ctrl(t, y)
        cost = (drift_s + drift_t) / diff + u_s - u_t
        rnd = rnd + 0.5 * (cost ** 2).sum(dim=-1, keepdim=True) * dt
        if lv:
            rnd = rnd + (cost * (sde_ctrl - u_s)).sum(dim=-1, keepdim=True) * dt
        rnd = rnd + (cost * db).sum(dim=-1, keepdim=True)
        x = y
        if return_traj:
            xs.append(x)
    rnd = rnd - terminal_unnorm_log_prob(x)
    return x, rnd, (torch.stack(xs) if return_traj else None)


def eubo_em(ts, x, noise, ctrl, sde, reference_ctrl, terminal_unnorm_log_prob, reference_log_prob,
            use_rescaling: bool = True):
    """EMReferenceSDELoss.compute_eubo, losses/oc.py:298-362.  ``use_rescaling`` is True for the EM loss and False
    for DDPMLikeReferenceSDELoss, which inherits this method (oc.py:571-582).  ``x`` is not mutated here (the
    reference mutates its input in place, lines 336-337)."""
    rnd = reference_log_prob(x).view((-1, 1)) - terminal_unnorm_log_prob(x)
    T = ts[-1]
    times_s = ts[:-1].flip((0,))
    times_t = ts[1:].flip((0,))
    mean_factors, var_factors = sde.transition_params(T - times_t, T - times_s)
    std_factors = var_factors.sqrt()
    for i, (s, t) in enumerate(zip(times_s, times_t)):
        z = noise[i]
        x = x * mean_factors[i] + std_factors[i] * z
        g = ctrl(T - s, x)
        ref = reference_ctrl(T - s, x)
        sde_diff = sde.diff(T - s)
        dt = t - s
        if use_rescaling:  # oc.py:345-346
            g = g / sde_diff
        running_cost = g * (ref + 0.5 * g)
        rnd = rnd - running_cost.sum(dim=-1, keepdim=True) * dt * sde_diff ** 2
        rnd = rnd + (g * x).sum(dim=-1, keepdim=True) * (1.0 / mean_factors[i] - 1.0 + sde.drift_coeff(T - s) * dt)
        rnd = rnd - (g * z).sum(dim=-1, keepdim=True) * (std_factors[i] / mean_factors[i])
    return rnd


def eubo_ei(ts, x, noise, ctrl, sde, reference_ctrl, terminal_unnorm_log_prob, reference_log_prob):
    """EIReferenceSDELoss.compute_eubo, losses/oc.py:512-568."""
    rnd = reference_log_prob(x).view((-1, 1)) - terminal_unnorm_log_prob(x)
    T = ts[-1]
    times_s = ts[:-1].flip((0,))
    times_t = ts[1:].flip((0,))
    mean_factors, var_factors = sde.transition_params(T - times_t, T - times_s)
    std_factors = var_factors.sqrt()
    for i, (s, t) in enumerate(zip(times_s, times_t)):
        z = noise[i]
        x = x * mean_factors[i] + std_factors[i] * z
        g = ctrl(T - s, x)
        ref = reference_ctrl(T - s, x)
        running_cost = g * (ref + 0.5 * g)
        rnd = rnd - running_cost.sum(dim=-1, keepdim=True) * sde.omega(s, t)
        rnd = rnd - (g * z).sum(dim=-1, keepdim=True) * torch.sqrt(sde.omega(s, t))
    return rnd


def eubo_cmcd(ts, x, noise, ctrl, drift, diff, terminal_unnorm_log_prob, initial_log_prob):
    """ControlledLangevinSDELoss.compute_eubo, losses/oc.py:757-828, including the reference quirk
    ``drift_s = sde.drift(t, y)`` (line 807: time t, not s)."""
    rnd = -terminal_unnorm_log_prob(x)
    times_s = ts[:-1].flip((0,))
    times_t = ts[1:].flip((0,))
    for i, (s, t) in enumerate(zip(times_s, times_t)):
        u_t = ctrl(t, x)
        dt = t - s
        db = dt.sqrt() * noise[i]
        drift_t = drift(t, x)
        y = x + (drift_t - u_t * diff) * dt + diff * db
        drift_s = drift(t, y)
        u_s = ctrl(s, y)
        cost = (drift_s + drift_t) / diff + u_s - u_t
        rnd = rnd - 0.5 * (cost ** 2).sum(dim=-1, keepdim=True) * dt
        rnd = rnd - (cost * db).sum(dim=-1, keepdim=True)
        x = y
    rnd = rnd + initial_log_prob(x)
    return rnd


# --------------------------------------------------------------------------------------
# estimators: losses/oc.py:134-173, eval/metrics.py:134-140, additions/hacking.py:14-33
# --------------------------------------------------------------------------------------


def compute_results(rnd: torch.Tensor, compute_weights: bool = True) -> dict:
    """BaseOCLoss.compute_results, losses/oc.py:134-173 plus ESS, eval/metrics.py:134-140."""
    out = {}
    neg = -rnd
    out["eval/elbo"] = neg.mean().item()
    if compute_weights:
        w = torch.softmax(neg, dim=0)
        out["log_norm_const_is"] = (neg.logsumexp(dim=0) - math.log(len(w))).item()
        out["eval/lv_loss"] = rnd.var().item()
        ess = (w.sum() ** 2 / (w ** 2).sum()).item()
        out["eval/effective_sample_size"] = ess
        out["eval/norm_effective_sample_size"] = ess / len(w)
    return out


def eubo_results(rnd_target: torch.Tensor) -> dict:
    """evaluate_eubo, sde_sampler/additions/hacking.py:22-32."""
    neg = -rnd_target
    w = torch.softmax(neg, dim=0)
    out = {
        "eval/log_norm_const_is_f": -rnd_target.logsumexp(dim=0).item() + math.log(len(w)),
        "eval/eubo": neg.mean().item(),
        "eval/effective_sample_size_f": (1.0 / (w ** 2).sum()).item(),
    }
    out["eval/norm_effective_sample_size_f"] = out["eval/effective_sample_size_f"] / len(w)
    return out


# --------------------------------------------------------------------------------------
# time grids: sde_sampler/utils/common.py:18-82
# --------------------------------------------------------------------------------------


def log_snr(sde, t):
    """OU.log_snr, sde_sampler/eq/sdes.py:347-351."""
    a = sde.s(t)
    return torch.log(torch.square(a) / (torch.square(a) * sde.sigma_sq(t)))


def get_timesteps(start, end, dt=None, steps=None, rescale_t=None, n_attemps=1024, sde=None):
    """get_timesteps, sde_sampler/utils/common.py:30-82 (uniform / cosine / snr bisection)."""
    if (steps is None) is (dt is None):
        raise ValueError("Exactly one of `dt` and `steps` should be defined.")
    if steps is None:
        steps = int(math.ceil((end - start) / dt))
    if sde is not None:
        f = lambda t: log_snr(sde, t)  # noqa: E731
        ls, le = f(torch.tensor(start)), f(torch.tensor(end))
        if torch.isnan(ls) or torch.isnan(le):
            raise ValueError("NaN SNR")
        rng = torch.linspace(ls, le, steps=steps + 1)
        low, high = start, end
        target = rng[1:-1]
        for _ in range(n_attemps):  # binary_search_v, common.py:18-27
            mid = (low + high) / 2.0
            ret = f(mid) if isinstance(mid, torch.Tensor) else f(torch.tensor(mid))
            low = torch.where(ret > target, mid, low)
            high = torch.where(ret <= target, mid, high)
        mid = (low + high) / 2.0
        return torch.concat([torch.FloatTensor([start]), mid, torch.FloatTensor([end])], dim=0).sort().values
    if rescale_t is None:
        return torch.linspace(start, end, steps=steps + 1)
    if rescale_t == "cosine":  # common.py:63-81
        s = 0.008
        pre_phase = torch.linspace(start, end, steps + 1) / end
        phase = ((pre_phase + s) / (1 + s)) * torch.pi * 0.5
        dts = torch.cos(phase) ** 4
        dts = dts / dts.sum()
        dts = dts * end
        return torch.concat((torch.tensor([start]), torch.cumsum(dts, -1)))
    raise ValueError("Unkown timestep rescaling method.")


# --------------------------------------------------------------------------------------
# one entry point over a neutral problem description (see tests/cases.py)
# --------------------------------------------------------------------------------------


def _cast(obj, dtype):
    if isinstance(obj, torch.Tensor):
        return obj.to(dtype) if obj.is_floating_point() else obj
    if isinstance(obj, dict):
        return {k: _cast(v, dtype) for k, v in obj.items()}
    return obj


def rollout(problem: dict, x0: torch.Tensor, noise: torch.Tensor, *, eubo: bool = False,
            compute_ito_int: bool = True, return_traj: bool = False, dtype=torch.float32, lv: bool = False):
    """Runs the rollout named by ``problem`` (keys: method, sde, ctrl, target, ref, ts, and for
    dds alpha/sigma, for cmcd diff/T/clip_score/prior).  Returns (x_T, rnd, xs) for the generative
    rollout, or rnd for ``eubo=True`` (x0 = target samples).  ``lv=True`` runs the TRAINING form of the linear
    losses (change_sde_ctrl=True, autograd through the control; see lv_loss_and_grads)."""
    if lv and eubo:
        raise ValueError("the lv training form belongs to the simulate loops")
    problem = _cast(problem, dtype)
    x0, noise = x0.to(dtype), noise.to(dtype)
    ts = problem["ts"]
    target_logp, target_score = make_target(problem["target"])
    if problem.get("clip_target") is not None:  # TrainableDiff.clipped_target_unnorm_log_prob, solver/oc.py:80-87
        raw = target_logp
        target_logp = lambda x: clip(raw(x), problem["clip_target"])  # noqa: E731
    ctrl = make_ctrl(problem["ctrl"], target_score, dtype)
    method = problem["method"]
    with (torch.enable_grad() if lv else torch.no_grad()):
        if method in ("em", "ei", "ddpm"):
            sde = make_sde(problem["sde"], dtype)
            ref_ctrl, ref_logp = make_reference(problem["ref"], sde)
            if eubo:
                if method == "ei":
                    return eubo_ei(ts, x0, noise, ctrl, sde, ref_ctrl, target_logp, ref_logp)
                return eubo_em(ts, x0, noise, ctrl, sde, ref_ctrl, target_logp, ref_logp, use_rescaling=(method == "em"))
            if method == "em":
                return simulate_em(ts, x0, noise, ctrl, sde, ref_ctrl, target_logp, ref_logp, return_traj, lv=lv)
            return simulate_ei(ts, x0, noise, ctrl, sde, ref_ctrl, target_logp, ref_logp, return_traj,
                               ddpm=(method == "ddpm"), lv=lv)
        if method == "dis_ei":  # the discrete-time DIS loss; prior as for DIS
            sde = make_sde(problem["sde"], dtype)
            _, prior_logp = make_reference(problem["ref"], None)
            if eubo:
                return eubo_dis_ei(ts, x0, noise, ctrl, sde, target_logp, prior_logp)
            return simulate_dis_ei(ts, x0, noise, ctrl, sde, target_logp, prior_logp, return_traj, lv=lv)
        if method == "dis":  # prior = IsotropicGauss (conf/prior/gauss.yaml), scale = sde.scale_diff_coeff (conf/solver/dis.yaml)
            sde = make_sde(problem["sde"], dtype)
            _, prior_logp = make_reference(problem["ref"], None)
            return simulate_dis(ts, x0, noise, ctrl, sde, target_logp, prior_logp, compute_ito_int, return_traj, lv=lv)
        if method == "dds":
            _, ref_logp = make_reference(problem["ref"], None)
            return simulate_dds(ts, x0, noise, ctrl, problem["alpha"], problem["sigma"], target_logp, ref_logp,
                                compute_ito_int, return_traj, lv=lv)
        if method == "cmcd":
            prior = problem["prior"]  # Gauss(loc, scale) diag; IsotropicGauss is the same arithmetic family
            ploc, pscale = prior["loc"], prior["scale"]
            if prior.get("isotropic", False):
                prior_logp = lambda x: isotropic_gauss_log_prob(x, float(ploc.flatten()[0]), float(pscale.flatten()[0]))  # noqa: E731
                prior_score = lambda x: (ploc.flatten()[0] - x) / pscale.flatten()[0] ** 2  # noqa: E731  gauss.py:764-766
            else:
                prior_logp = lambda x: gmm_log_prob(x, ploc.reshape(1, -1), pscale.reshape(1, -1), None)  # noqa: E731
                prior_score = lambda x: score_gauss(x, ploc.reshape(1, -1), torch.square(pscale.reshape(1, -1)))  # noqa: E731
            diff = torch.tensor(problem["diff"], dtype=dtype)
            T = torch.tensor(problem["T"], dtype=dtype)
            drift = lambda t, x: langevin_drift(t, x, target_score, prior_score, diff, T, problem.get("clip_score"))  # noqa: E731
            if eubo:
                return eubo_cmcd(ts, x0, noise, ctrl, drift, diff, target_logp, prior_logp)
            return simulate_cmcd(ts, x0, noise, ctrl, drift, diff, target_logp, prior_logp, return_traj, lv=lv)
    raise ValueError(method)


def lv_loss_and_grads(problem: dict, x0: torch.Tensor, noise: torch.Tensor, max_rnd: float | None = 1e8,
                      dtype=torch.float32, traj_per_sample: int = 1):
    """The training objective ``loss(ts, x, ...)`` of the linear rollout losses with method='lv' and its gradient:
    __call__ (losses/oc.py:364-394, 1240-1272, 1399-1431: change_sde_ctrl=True, compute_ito_int=True) ->
    BaseOCLoss.compute_loss (105-131: ``rnd[mask].var()`` over the particles that pass ``filter``, 67-81; with
    ``traj_per_sample`` > 1 the 'lv_traj' form: initial values repeated, 383-384, variance over each sample's trajectories,
    118-124).  Returns (loss, {state_dict key: d loss / d parameter}, rnd)."""
    if traj_per_sample != 1:
        x0 = x0.repeat(traj_per_sample, 1, 1).reshape(-1, x0.shape[-1])
    problem = dict(problem)
    ctrl = dict(problem["ctrl"])
    ctrl["sd"] = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in ctrl["sd"].items()}
    problem["ctrl"] = ctrl
    cast = _cast({k: v for k, v in problem.items() if k != "ctrl"}, dtype)
    cast["ctrl"] = ctrl
    _, rnd, _ = rollout(cast, x0, noise, compute_ito_int=True, dtype=dtype, lv=True)
    mask = rnd.isfinite() if max_rnd is None else rnd < max_rnd
    if traj_per_sample != 1:
        loss = rnd.reshape(traj_per_sample, -1, 1)[:, mask.reshape(traj_per_sample, -1, 1).all(dim=0)].var(dim=0).mean()
    else:
        loss = rnd[mask].var()
    names = list(ctrl["sd"])
    grads = torch.autograd.grad(loss, [ctrl["sd"][k] for k in names], allow_unused=True)
    return loss.detach(), {k: (g if g is not None else torch.zeros_like(ctrl["sd"][k])) for k, g in zip(names, grads)}, rnd.detach()


# --------------------------------------------------------------------------------------
# MCMC: sde_sampler/additions/mcmc.py, experiments/benchmark_utils.py:268-333
# --------------------------------------------------------------------------------------


def heuristics_step_size(stepsize, log_acc, target_acceptance=0.75, factor=1.01, tol=0.05):
    """additions/mcmc.py:54-72 (stepsize (C, 1), log_acc (C,))."""
    up = (log_acc - math.log(target_acceptance) > math.log1p(tol)).view(-1, 1)
    stepsize = torch.where(up, stepsize * factor, stepsize)
    down = (math.log(target_acceptance) - log_acc > -math.log1p(-tol)).view(-1, 1)
    return torch.where(down, stepsize / factor, stepsize)


def mala_chains(target: dict, y_init, step_size: float, n_warmup: int, n_steps: int, noise, unif, adapt=True,
                dtype=torch.float32):
    """The loop of mcmc_sample(mcmc_type='mala') (experiments/benchmark_utils.py:268-333) around mala_step
    (additions/mcmc.py:75-134) on recorded draws: noise [S, C, d] for the proposals, unif [S, C] for the accept test.
    The reference differentiates unnorm_log_prob by autograd; the score restated above is the same function.
    Returns (ys [n_steps, C, d], step_size (C, 1), log_acc [S, C])."""
    logp_fn, score_fn = make_target(_cast(target, dtype))
    y = y_init.to(dtype).clone()
    h = step_size * torch.ones((y.shape[0], 1), dtype=dtype)
    logp, grad = logp_fn(y).flatten(), score_fn(y)
    ys, accs = [], []
    with torch.no_grad():
        for k in range(n_warmup + n_steps):
            mean = y + h * grad
            var = 2.0 * h
            y_prop = torch.sqrt(var) * noise[k].to(dtype) + mean
            logp_p, grad_p = logp_fn(y_prop).flatten(), score_fn(y_prop)
            joint_prop = logp_p - (-0.5 * torch.sum(torch.square(y_prop - mean), dim=-1)) / var.flatten()
            joint_orig = logp - (-0.5 * torch.sum(torch.square(y - (y_prop + h * grad_p)), dim=-1)) / var.flatten()
            log_acc = joint_prop - joint_orig
            mask = torch.log(unif[k].to(dtype)) < log_acc
            y = torch.where(mask.view(-1, 1), y_prop, y)
            grad = torch.where(mask.view(-1, 1), grad_p, grad)
            logp = torch.where(mask, logp_p, logp)
            if adapt:
                h = heuristics_step_size(h, log_acc)
            accs.append(log_acc)
            if k >= n_warmup:
                ys.append(y.clone())
    return torch.stack(ys), h, torch.stack(accs)


def rwmh_chains(target: dict, y_init, step_size: float, n_warmup: int, n_steps: int, noise, unif, adapt=True,
                dtype=torch.float32):
    """The same loop with mcmc_type='rwmh': rwmh_step (additions/mcmc.py:258-290) - proposal y + step_size z, acceptance
    log u < log p(y') - log p(y) - on recorded draws.  Returns (ys [n_steps, C, d], step_size (C, 1), log_acc [S, C])."""
    logp_fn, _ = make_target(_cast(target, dtype))
    y = y_init.to(dtype).clone()
    h = step_size * torch.ones((y.shape[0], 1), dtype=dtype)
    logp = logp_fn(y).flatten()
    ys, accs = [], []
    with torch.no_grad():
        for k in range(n_warmup + n_steps):
            y_prop = y + h * noise[k].to(dtype)
            logp_p = logp_fn(y_prop).flatten()
            log_acc = logp_p - logp
            mask = torch.log(unif[k].to(dtype)) < log_acc
            y = torch.where(mask.view(-1, 1), y_prop, y)
            logp = torch.where(mask, logp_p, logp)
            if adapt:
                h = heuristics_step_size(h, log_acc)
            accs.append(log_acc)
            if k >= n_warmup:
                ys.append(y.clone())
    return torch.stack(ys), h, torch.stack(accs)

