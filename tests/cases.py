"""Neutral, seed-driven descriptions of the rollout problems used by the parity tests.

A *problem* is a plain dict of python scalars and float32 torch tensors (no classes) that three
independent builders consume: the reference (``oracle/make_golden.py``, build container only),
the oracle (``oracle.rollout_oracle.rollout``) and the product (``tests/product_builders.py``).
Shapes follow BASELINE.json's configs; batches are small so the CPU oracle finishes
in seconds, and B is deliberately not a multiple of the 128-particle tile.

Weights: every Linear is drawn U(+-1/sqrt(fan_in)) like nn.Linear's default, INCLUDING the last
layers (the reference zero-inits those at scale 1e-6, models/utils.py:7-22, which would make the
control ~0 and the test vacuous); ``out_gain``/``gamma`` scale the control so that trajectories
stay in a numerically sane range.
"""
from __future__ import annotations

import math

import torch

C = 64  # FourierMLP / TimeEmbed channels (conf/model/base/fouriermlp.yaml, time_embed.yaml)


def _lin(g, out_f, in_f, gain=1.0):
    b = gain / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * b
    bias = (torch.rand(out_f, generator=g) * 2 - 1) * b
    return w, bias


def ctrl_state_dict(d: int, kind: str, seed: int, out_gain: float = 1.0, gamma: float = 0.3,
                    num_hidden: int = 2) -> dict:
    """state_dict with the key names of ScoreCtrl/ClippedCtrl(FourierMLP[, TimeEmbed]) in the reference."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def put(name, w, b):
        sd[name + ".weight"], sd[name + ".bias"] = w, b
    put("base_model.input_embed", *_lin(g, C, d))
    sd["base_model.timestep_embed.timestep_phase"] = torch.randn(1, C, generator=g)
    put("base_model.timestep_embed.hidden_layer.0", *_lin(g, C, 2 * C))
    put("base_model.timestep_embed.out_layer", *_lin(g, C, C))
    for i in range(num_hidden):
        put(f"base_model.hidden_layer.{i}", *_lin(g, C, C))
    put("base_model.out_layer", *_lin(g, d, C, gain=out_gain))
    if kind == "score":
        sd["score_model.timestep_phase"] = torch.randn(1, C, generator=g)
        put("score_model.hidden_layer.0", *_lin(g, C, 2 * C))
        put("score_model.hidden_layer.1", *_lin(g, C, C))
        put("score_model.hidden_layer.2", *_lin(g, C, C))
        w, b = _lin(g, 1, C, gain=gamma)
        put("score_model.out_layer", w, b + gamma)
    return sd


def ctrl(d, kind, seed, sde=None, prior=None, **kw):
    """kind: 'clipped' (ClippedCtrl), 'score' (ScoreCtrl), and the two DIS parametrisations 'cancel' (CancelDriftCtrl,
    needs ``sde``) and 'lerp' (LerpCtrl, needs ``sde`` and the IsotropicGauss ``prior`` {'loc', 'scale'})."""
    out = {"kind": kind, "sd": ctrl_state_dict(d, "clipped" if kind == "clipped" else "score", seed, **kw),
           "clip_model": 1e4, "clip_score": None if kind == "clipped" else 1e4, "scale_score": 1.0}
    if kind in ("cancel", "lerp"):
        out["sde"] = dict(sde)
    if kind == "lerp":
        out["prior"] = dict(prior)
    return out


# ---- targets (parameters restated from the reference constructors) --------------------------


def two_modes(dim=2, a=1.0, ill="medium"):
    """TwoModes, distr/gauss.py:422-453."""
    w = torch.tensor([2.0, 1.0])
    loc = torch.stack([-a * torch.ones(dim), a * torch.ones(dim)])
    if ill == "medium":
        scale = torch.sqrt(0.05 * torch.logspace(-1, 0.0, dim)).unsqueeze(0).expand(2, -1).contiguous()
    elif ill == "hard":
        scale = torch.sqrt(0.05 * torch.logspace(-2.0, 0.0, dim)).unsqueeze(0).expand(2, -1).contiguous()
    else:
        scale = torch.sqrt(0.05 * torch.ones_like(loc))
    return {"kind": "gmm", "loc": loc, "scale": scale, "weights": w}


def many_modes(n_modes=16, dim=50, seed_loc=42, factor=3.0, var=0.5):
    """ManyModes, distr/gauss.py:569-594."""
    g = torch.Generator().manual_seed(seed_loc)
    w = torch.logspace(0.0, 1.0, n_modes, base=factor)
    loc = 2 * n_modes * torch.rand((n_modes, dim), generator=g) - n_modes
    scale = torch.sqrt(var * torch.ones_like(loc))
    return {"kind": "gmm", "loc": loc, "scale": scale, "weights": w}


def phi4(dim=100, a=0.1, b=0.0, beta=20.0):
    return {"kind": "phi4", "a": a, "b": b, "beta": beta, "dim": dim}


def logreg_synthetic(n=166, p=60, seed=7, weight_scale=4.5, intercept_mean=-2.5, intercept_scale=0.5):
    """Synthetic-shape twin of data/sonar.pkl (166x60) / ionosphere.pkl (280x33): X~U(0,1), y~Bern(1/2)
    (BASELINE.json config 4); prior scales from conf/target/{sonar,ionosphere}.yaml."""
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n, p, generator=g)
    y = (torch.rand(n, generator=g) < 0.5).float()
    return {"kind": "logreg", "X": X, "y": y, "weight_scale": weight_scale,
            "intercept_mean": intercept_mean, "intercept_scale": intercept_scale, "dim": p + 1}


VP10 = {"kind": "vp", "beta_min": 0.1, "beta_max": 10.0, "scale": 1.0, "T": 1.0}
VP20 = {"kind": "vp", "beta_min": 0.1, "beta_max": 20.0, "scale": 1.0, "T": 1.0}
PBM = {"kind": "pbm", "diff": 0.4472135954999579, "T": 5.0}
BM = {"kind": "bm", "diff": 0.4472135954999579, "T": 5.0}


def uniform_ts(T, K, start=0.0):
    return torch.linspace(start, T, K + 1)


def cosine_ts(end=6.4, dt=0.05):
    """get_timesteps(rescale_t='cosine'), utils/common.py:63-81, as DDS uses it (conf/solver/dds.yaml)."""
    steps = int(math.ceil(end / dt))
    s = 0.008
    pre = torch.linspace(0.0, end, steps + 1) / end
    dts = torch.cos(((pre + s) / (1 + s)) * torch.pi * 0.5) ** 4
    dts = dts / dts.sum() * end
    return torch.concat((torch.tensor([0.0]), torch.cumsum(dts, -1)))


def snr_ts_vp(sde: dict, K: int):
    """snr-uniform grid for VP (utils/common.py:44-56): start 1e-4, end T-1e-4 (benchmark_utils.py:188-189).
    Computed by 1024-step bisection in float32 exactly like the reference formulas."""
    bmin, bmax, T, c = (torch.tensor(sde[k]) for k in ("beta_min", "beta_max", "T", "scale"))

    def log_snr(t):
        a = torch.exp(-0.5 * (bmin * t + (0.5 * t ** 2 / T) * (bmax - bmin)))
        return torch.log(torch.square(a) / (torch.square(a) * (-c ** 2 * (1.0 - (1.0 / a ** 2)))))
    start, end = 1e-4, float(T) - 1e-4
    rng = torch.linspace(log_snr(torch.tensor(start)), log_snr(torch.tensor(end)), steps=K + 1)[1:-1]
    low, high = start, end
    for _ in range(1024):
        mid = (low + high) / 2.0
        ret = log_snr(mid if isinstance(mid, torch.Tensor) else torch.tensor(mid))
        low = torch.where(ret > rng, mid, low)
        high = torch.where(ret <= rng, mid, high)
    mid = (low + high) / 2.0
    return torch.concat([torch.FloatTensor([start]), mid, torch.FloatTensor([end])]).sort().values


# ---- small public interfaces: time grids and the Euler integrator (goldens: oracle/make_golden.py --aux) ----------
AUX_SDES = {"vp10": VP10, "vp20": VP20, "vpcos": {"kind": "vpcos", "c": 0.008, "scale": 1.0, "T": 1.0}, "pbm": PBM}
AUX_GRIDS = {  # name -> (sde or None, get_timesteps keyword arguments)
    "uniform_steps": (None, dict(start=0.0, end=1.0, steps=37)),
    "uniform_dt": (None, dict(start=1e-3, end=5.0, dt=0.07)),
    "cosine_dds": (None, dict(start=0.0, end=6.4, dt=0.05, rescale_t="cosine")),
    "cosine_steps": (None, dict(start=0.0, end=3.2, steps=41, rescale_t="cosine")),
    "snr_vp10": ("vp10", dict(start=1e-4, end=1.0 - 1e-4, steps=100)),
    "snr_vp20": ("vp20", dict(start=1e-4, end=1.0 - 1e-4, steps=200)),
    "snr_vp20_short": ("vp20", dict(start=1e-4, end=1.0 - 1e-4, steps=17)),
    "snr_vpcos": ("vpcos", dict(start=1e-3, end=1.0 - 1e-3, steps=64)),
    "snr_pbm": ("pbm", dict(start=1e-4, end=5.0 - 1e-4, steps=100)),
}


def euler_case():
    """EulerIntegrator over the uncontrolled VP SDE: 40 Euler steps of dt = 0.025, output times partly off the grid."""
    g = torch.Generator().manual_seed(77)
    B, d = 33, 5
    return {"sde": VP10, "dt": 0.025, "x0": torch.randn(B, d, generator=g), "noise": torch.randn(64, B, d, generator=g),
            "ts": torch.tensor([0.0, 0.025, 0.06, 0.31, 0.5, 0.777, 1.0]),
            "snr_steps": 24, "ts_snr": torch.tensor([1e-4, 0.2, 0.5, 0.9, 1.0 - 1e-4])}


# ---- cases ---------------------------------------------------------------------------------


def case_em_two_modes(ctrl_kind="score"):
    """BASELINE config 1: TwoModes 2-D, RDS vp-ref (VP 0.1..10), gaussian ref, EM, K=100."""
    d = 2
    return {
        "problem": {"method": "em", "sde": VP10, "ts": uniform_ts(1.0, 100), "target": two_modes(d),
                    "ctrl": ctrl(d, ctrl_kind, seed=11, out_gain=1.0, gamma=0.2),
                    "ref": {"kind": "gauss", "mean": torch.zeros(d), "var": 1.5 * torch.ones(d)}},
        "B": 200, "seed": 101, "prior": ("iso", 0.0, 1.0)}


def case_ei_many_modes(K=200, B=200, d=50, M=16):
    """BASELINE config 2: ManyModes d=50 (16 modes), RDS vp-ref beta_max=20, GMM ref, EI, K=200, ScoreCtrl."""
    tgt = many_modes(M, d)
    ref = {"kind": "gmm", "means": tgt["loc"] + 0.1, "variances": 1.2 * tgt["scale"] ** 2,
           "weights": tgt["weights"].clone()}
    return {
        "problem": {"method": "ei", "sde": VP20, "ts": uniform_ts(1.0, K), "target": tgt,
                    "ctrl": ctrl(d, "score", seed=12, out_gain=1.0, gamma=0.05), "ref": ref},
        "B": B, "seed": 102, "prior": ("iso", 0.0, 1.0)}


def case_ei_close_modes(method="ei"):
    """Shared-variance mixtures whose modes overlap and sit far from the origin: near-ties of the responsibilities are
    common and the logits are large, so the mixture kernel's logit GEMM (lrds_rollout_mix.cuh) must hand many particles
    over to the exact quadratic forms while others keep its result."""
    d, M = 24, 10
    g = torch.Generator().manual_seed(5)
    loc = 8.0 + 1.2 * (torch.rand(M, d, generator=g) - 0.5)
    tgt = {"kind": "gmm", "loc": loc, "scale": math.sqrt(0.5) * torch.ones(M, d), "weights": torch.logspace(0.0, 1.0, M, base=2.0)}
    ref = {"kind": "gmm", "means": loc + 0.1, "variances": 0.6 * torch.ones(M, d), "weights": tgt["weights"].clone()}
    return {
        "problem": {"method": method, "sde": VP10, "ts": uniform_ts(1.0, 60), "target": tgt,
                    "ctrl": ctrl(d, "score", seed=33, out_gain=0.5, gamma=0.03), "ref": ref},
        "B": 130, "seed": 133, "prior": ("iso", 0.0, 1.0)}


def case_ei_two_modes_gauss():
    """RDS vp-ref with a Gaussian reference and the EI integrator over a mixture target (cfg 1's solver with cfg 2's
    integrator): target score on the tensor core, reference score one FMA per dim."""
    d = 7
    return {
        "problem": {"method": "ei", "sde": VP10, "ts": uniform_ts(1.0, 60), "target": two_modes(d),
                    "ctrl": ctrl(d, "score", seed=21, out_gain=0.5, gamma=0.002),
                    "ref": {"kind": "gauss", "mean": 0.1 * torch.ones(d), "var": 1.5 * torch.ones(d)}},
        "B": 150, "seed": 111, "prior": ("iso", 0.0, 1.0)}


def case_ei_phi4_gmm():
    """RDS with a mixture reference over the PhiFour lattice (the reference's experiments/sample_phi_four_gmm_mcmc.py):
    lattice stencil for the target score, reference contraction on the tensor core."""
    d = 20
    g = torch.Generator().manual_seed(9)
    means = torch.stack([torch.ones(d), -torch.ones(d), 0.2 * torch.randn(d, generator=g)])
    ref = {"kind": "gmm", "means": means, "variances": 0.3 + 0.2 * torch.rand(3, d, generator=g),
           "weights": torch.tensor([2.0, 2.0, 1.0])}
    return {
        "problem": {"method": "ei", "sde": VP10, "ts": uniform_ts(1.0, 80), "target": phi4(d),
                    "ctrl": ctrl(d, "score", seed=22, out_gain=0.3, gamma=0.004), "ref": ref},
        "B": 120, "seed": 112, "prior": ("iso", 0.0, 1.0)}


def case_ei_many_modes_clipped():
    """RDS with a mixture reference and the plain ClippedCtrl drift (base_zero_init): no target score."""
    d, M = 12, 6
    tgt = many_modes(M, d)
    ref = {"kind": "gmm", "means": tgt["loc"] - 0.2, "variances": 1.5 * tgt["scale"] ** 2, "weights": torch.ones(M)}
    return {
        "problem": {"method": "ei", "sde": VP20, "ts": uniform_ts(1.0, 70), "target": tgt,
                    "ctrl": ctrl(d, "clipped", seed=23, out_gain=1.0), "ref": ref},
        "B": 100, "seed": 113, "prior": ("iso", 0.0, 1.0)}


def case_em_many_modes_gmm():
    """RDS with a mixture reference and the Euler-Maruyama integrator (integrator_type='em' with ref_type='gmm')."""
    d, M = 9, 5
    tgt = many_modes(M, d)
    ref = {"kind": "gmm", "means": tgt["loc"] + 0.1, "variances": 1.3 * tgt["scale"] ** 2, "weights": tgt["weights"].clone()}
    return {
        "problem": {"method": "em", "sde": VP10, "ts": uniform_ts(1.0, 120), "target": tgt,
                    "ctrl": ctrl(d, "score", seed=24, out_gain=1.0, gamma=0.05), "ref": ref},
        "B": 110, "seed": 114, "prior": ("iso", 0.0, 1.0)}


def case_pis_many_modes():
    """PIS over a mixture target (pis_orig of experiments/sample_many_modes_competing.py)."""
    d, M = 10, 7
    return {
        "problem": {"method": "em", "sde": BM, "ts": uniform_ts(5.0, 90), "target": many_modes(M, d),
                    "ctrl": ctrl(d, "score", seed=27, out_gain=0.5, gamma=0.02),
                    "ref": {"kind": "pis", "loc": torch.zeros(d)}},
        "B": 120, "seed": 117, "prior": ("delta", 0.0)}


def case_dds_many_modes(compute_ito_int=True):
    """DDS over a mixture target (dds_orig)."""
    d, M = 10, 7
    return {
        "problem": {"method": "dds", "sde": None, "alpha": 1.0, "sigma": 2.0, "ts": cosine_ts(6.4, 0.08),
                    "target": many_modes(M, d), "ctrl": ctrl(d, "score", seed=28, out_gain=0.5, gamma=0.02),
                    "ref": {"kind": "iso", "loc": 0.0, "scale": 2.0}},
        "B": 120, "seed": 118, "prior": ("iso", 0.0, 2.0), "compute_ito_int": compute_ito_int}


def case_dis(target="many_modes", compute_ito_int=True, scale=1.0, ctrl_kind="score"):
    """DIS (dis_orig: solver Bridge + TimeReversalLoss with inference_ctrl=None, conf/solver/dis.yaml): VP 0.1..10,
    Euler-Maruyama, control at the loop time, prior IsotropicGauss(scale=sde.scale_diff_coeff); drift models ScoreCtrl
    (target_informed_zero_init), CancelDriftCtrl (target_informed_langevin_init) or LerpCtrl (the solver's default,
    target_informed_lerp_tempering)."""
    sde = dict(VP10, scale=scale)
    prior = {"loc": 0.0, "scale": scale}
    if target == "many_modes":
        d, K, B, tgt, c = 10, 90, 120, many_modes(7, 10), dict(seed=31, out_gain=0.5, gamma=0.02)
    elif target == "phi4":
        d, K, B, tgt, c = 24, 80, 100, phi4(24), dict(seed=32, out_gain=0.3, gamma=0.004)
    else:
        d, K, B, tgt, c = 61, 40, 96, logreg_synthetic(166, 60), dict(seed=33, out_gain=0.5, gamma=0.01)
    return {
        "problem": {"method": "dis", "sde": sde, "ts": uniform_ts(1.0, K), "target": tgt,
                    "ctrl": ctrl(d, ctrl_kind, sde=sde, prior=prior, **c), "ref": {"kind": "iso", **prior}},
        "B": B, "seed": 120 + len(target), "prior": ("iso", 0.0, scale), "compute_ito_int": compute_ito_int}


def case_dis_ei(target="many_modes"):
    """Discrete-time DIS loss (DiscreteTimeReversalLossEI, losses/oc.py:897-1103): exponential integrator over VP, no
    reference control, the log-weight starting at the prior log-density."""
    case = case_dis(target, True)
    p = case["problem"]
    p["method"] = "dis_ei"
    # the time-reversed exponential-integrator step expands x by sqrt(1 + lambda) per step (12x over the horizon) unless the
    # control holds it: a score factor of the right size (gamma * target score ~ -(x - mu)) keeps the log-weights at
    # O(10^2), where the absolute 1e-3 bar on log Z is above the float32 resolution
    d = 10 if target == "many_modes" else 24
    p["ctrl"] = ctrl(d, "score", seed=34, out_gain=0.3, gamma=0.8 if target == "many_modes" else 0.04)
    case["seed"] += 40
    return case


VPCOS = {"kind": "vpcos", "c": 0.008, "scale": 1.0, "T": 1.0}


def case_ei_cosine(method="ei"):
    """RDS vp-ref with the cosine schedule (make_model(force_vp_cosine=True): conf/sde/vp_cos.yaml, grid start 1e-3,
    experiments/benchmark_utils.py:172-173, 191-192), mixture reference, uniform grid."""
    d, M = 9, 5
    tgt = many_modes(M, d)
    ref = {"kind": "gmm", "means": tgt["loc"] + 0.1, "variances": 1.3 * tgt["scale"] ** 2, "weights": tgt["weights"].clone()}
    return {
        "problem": {"method": method, "sde": VPCOS, "ts": uniform_ts(1.0, 300, start=1e-3), "target": tgt,
                    "ctrl": ctrl(d, "score", seed=35, out_gain=1.0, gamma=0.05), "ref": ref},
        "B": 110, "seed": 135, "prior": ("iso", 0.0, 1.0)}


def case_ei_phi4_gauss():
    """RDS vp-ref with its default Gaussian reference over the PhiFour lattice (experiments/sample_phi_four_competing.py)."""
    d = 24
    return {
        "problem": {"method": "ei", "sde": VP10, "ts": uniform_ts(1.0, 80), "target": phi4(d),
                    "ctrl": ctrl(d, "score", seed=29, out_gain=0.3, gamma=0.004),
                    "ref": {"kind": "gauss", "mean": torch.zeros(d), "var": 1.2 * torch.ones(d)}},
        "B": 100, "seed": 119, "prior": ("iso", 0.0, 1.0)}


def case_ddpm_snr():
    """DDPM-like integrator on an snr grid (API parity), TwoModes d=5, GMM ref."""
    d = 5
    tgt = two_modes(d)
    ref = {"kind": "gmm", "means": tgt["loc"] * 0.9, "variances": 2.0 * tgt["scale"] ** 2,
           "weights": torch.tensor([1.0, 1.0])}
    return {
        "problem": {"method": "ddpm", "sde": VP10, "ts": snr_ts_vp(VP10, 50), "target": tgt,
                    "ctrl": ctrl(d, "score", seed=13, out_gain=0.5, gamma=0.002), "ref": ref},
        "B": 130, "seed": 103, "prior": ("iso", 0.0, 1.0)}


def case_ei_pbm():
    """pbm-ref: PinnedBM, Delta prior, start 1e-4 (conf/solver/pbm_rds.yaml), default reference
    (solver/oc.py:539-545: x_init = prior.loc, var_init = T sigma^2), EI, ManyModes d=8 M=4, ClippedCtrl."""
    d = 8
    return {
        "problem": {"method": "ei", "sde": PBM, "ts": uniform_ts(5.0 - 1e-4, 64, start=1e-4),
                    "target": many_modes(4, d, var=0.5), "ctrl": ctrl(d, "clipped", seed=14, out_gain=0.5),
                    "ref": {"kind": "gauss", "mean": torch.zeros(d), "var": 5.0 * 0.4472135954999579 ** 2 * torch.ones(d)}},
        "B": 150, "seed": 104, "prior": ("delta", 0.0)}


def case_pis_phi4(K=256, B=64):
    """BASELINE config 3 (PIS): PhiFour d=100, ScaledBM sigma=sqrt(.2) T=5, Delta prior, EM, K=256."""
    d = 100
    return {
        "problem": {"method": "em", "sde": BM, "ts": uniform_ts(5.0, K), "target": phi4(d),
                    "ctrl": ctrl(d, "score", seed=15, out_gain=0.3, gamma=0.004),
                    "ref": {"kind": "pis", "loc": torch.zeros(d)}},
        "B": B, "seed": 105, "prior": ("delta", 0.0)}


def case_dds_phi4(compute_ito_int=True, B=64):
    """BASELINE config 3 (DDS): PhiFour d=100, alpha=sigma=1, cosine grid dt=.05 end 6.4 (K=128)."""
    d = 100
    return {
        "problem": {"method": "dds", "sde": None, "alpha": 1.0, "sigma": 1.0, "ts": cosine_ts(6.4, 0.05),
                    "target": phi4(d), "ctrl": ctrl(d, "score", seed=16, out_gain=0.3, gamma=0.004),
                    "ref": {"kind": "iso", "loc": 0.0, "scale": 1.0}},
        "B": B, "seed": 106, "prior": ("iso", 0.0, 1.0), "compute_ito_int": compute_ito_int}


def case_cmcd_logreg(n=166, p=60, K=32, B=96, ctrl_kind="score"):
    """BASELINE config 4: CMCD, sigma=1 T=1 clip 1e5, prior N(0, 5^2 I), logistic regression (sonar shape)."""
    tgt = logreg_synthetic(n, p)
    d = p + 1
    return {
        "problem": {"method": "cmcd", "sde": None, "diff": 1.0, "T": 1.0, "clip_score": 1e5,
                    "ts": uniform_ts(1.0, K), "target": tgt,
                    "ctrl": ctrl(d, ctrl_kind, seed=17, out_gain=0.5, gamma=0.02),
                    "prior": {"loc": torch.zeros(d), "scale": 5.0 * torch.ones(d), "isotropic": True}},
        "B": B, "seed": 107, "prior": ("iso", 0.0, 5.0)}


def case_pis_logreg(n=166, p=60, K=40, B=96):
    """PIS over the logistic-regression posterior (experiments/sample_bayesian_logreg_competing.py, pis_orig)."""
    d = p + 1
    return {
        "problem": {"method": "em", "sde": BM, "ts": uniform_ts(5.0, K), "target": logreg_synthetic(n, p),
                    "ctrl": ctrl(d, "score", seed=25, out_gain=0.5, gamma=0.01),
                    "ref": {"kind": "pis", "loc": torch.zeros(d)}},
        "B": B, "seed": 115, "prior": ("delta", 0.0)}


def case_dds_logreg(n=280, p=33, B=96):
    """DDS over the logistic-regression posterior (dds_orig; cosine grid), ionosphere shape."""
    d = p + 1
    return {
        "problem": {"method": "dds", "sde": None, "alpha": 1.0, "sigma": 1.0, "ts": cosine_ts(6.4, 0.1),
                    "target": logreg_synthetic(n, p), "ctrl": ctrl(d, "score", seed=26, out_gain=0.3, gamma=0.01),
                    "ref": {"kind": "iso", "loc": 0.0, "scale": 1.0}},
        "B": B, "seed": 116, "prior": ("iso", 0.0, 1.0), "compute_ito_int": True}


def case_cmcd_gmm():
    """CMCD on a GMM target with a fitted diagonal Gaussian prior (CMCD.update_prior, solver/oc.py:291-303)."""
    d = 8
    tgt = many_modes(4, d, var=0.5)
    g = torch.Generator().manual_seed(5)
    return {
        "problem": {"method": "cmcd", "sde": None, "diff": 1.0, "T": 1.0, "clip_score": 1e5,
                    "ts": uniform_ts(1.0, 48), "target": tgt, "ctrl": ctrl(d, "clipped", seed=18, out_gain=0.5),
                    "prior": {"loc": 0.3 * torch.randn(d, generator=g), "scale": 2.0 + torch.rand(d, generator=g),
                              "isotropic": False}},
        "B": 140, "seed": 108, "prior": ("gauss",)}


def case_cmcd_many_modes(ctrl_kind="score"):
    """CMCD over a ManyModes target with the target-informed drift model (experiments/sample_many_modes_competing.py:100-115;
    isotropic Gaussian prior, conf/solver/cmcd.yaml): the CMCD mixture kernel (lrds_rollout_cmcd_mix.cuh), odd number of
    8-dim chunks, a batch that does not fill its last tile."""
    d, M = 20, 6
    return {
        "problem": {"method": "cmcd", "sde": None, "diff": 1.0, "T": 1.0, "clip_score": 1e5,
                    "ts": uniform_ts(1.0, 40), "target": many_modes(M, d, var=0.5),
                    "ctrl": ctrl(d, ctrl_kind, seed=41, out_gain=0.5, gamma=0.02),
                    "prior": {"loc": torch.zeros(d), "scale": 5.0 * torch.ones(d), "isotropic": True}},
        "B": 170, "seed": 141, "prior": ("iso", 0.0, 5.0)}


def _lv_traj(case, traj_per_sample=2):
    """method='lv_traj': every initial value is repeated traj_per_sample times (losses/oc.py:383-384) and the loss is the
    mean over the samples of the variance over their trajectories (118-124).  ``noise_for`` draws for B * traj particles."""
    case = dict(case)
    case["traj_per_sample"] = traj_per_sample
    return case


def _eubo(case, seed):
    case = dict(case)
    case["eubo"] = True
    case["seed"] = seed
    return case


CASES = {
    "em_two_modes_score": lambda: case_em_two_modes("score"),
    "em_two_modes_clipped": lambda: case_em_two_modes("clipped"),
    "ei_many_modes": lambda: case_ei_many_modes(),
    "ddpm_snr": case_ddpm_snr,
    "ei_close_modes": case_ei_close_modes,
    "em_close_modes": lambda: case_ei_close_modes("em"),
    "eubo_ei_close_modes": lambda: _eubo(case_ei_close_modes(), 209),
    "ei_two_modes_gauss": case_ei_two_modes_gauss,
    "ei_phi4_gmm": case_ei_phi4_gmm,
    "ei_many_modes_clipped": case_ei_many_modes_clipped,
    "em_many_modes_gmm": case_em_many_modes_gmm,
    "pis_many_modes": case_pis_many_modes,
    "dds_many_modes_ito": lambda: case_dds_many_modes(True),
    "dds_many_modes_noito": lambda: case_dds_many_modes(False),
    "ei_phi4_gauss": case_ei_phi4_gauss,
    "ei_pbm": case_ei_pbm,
    "pis_phi4": lambda: case_pis_phi4(),
    "dds_phi4_ito": lambda: case_dds_phi4(True),
    "dds_phi4_noito": lambda: case_dds_phi4(False),
    "cmcd_logreg_sonar": lambda: case_cmcd_logreg(166, 60),
    "cmcd_logreg_iono": lambda: case_cmcd_logreg(280, 33, ctrl_kind="clipped"),
    "cmcd_gmm": case_cmcd_gmm,
    "cmcd_many_modes_score": case_cmcd_many_modes,
    "pis_logreg": lambda: case_pis_logreg(),
    "dds_logreg": lambda: case_dds_logreg(),
    "dis_many_modes_ito": lambda: case_dis("many_modes", True),
    "dis_many_modes_noito": lambda: case_dis("many_modes", False, scale=1.5),
    "dis_phi4": lambda: case_dis("phi4"),
    "dis_logreg": lambda: case_dis("logreg"),
    "dis_many_modes_lerp": lambda: case_dis("many_modes", ctrl_kind="lerp", scale=1.25),
    "dis_many_modes_langevin": lambda: case_dis("many_modes", ctrl_kind="cancel"),
    "dis_phi4_langevin": lambda: case_dis("phi4", False, ctrl_kind="cancel"),
    "dis_logreg_lerp": lambda: case_dis("logreg", ctrl_kind="lerp"),
    "dis_ei_many_modes": lambda: case_dis_ei("many_modes"),
    "dis_ei_phi4": lambda: case_dis_ei("phi4"),
    "eubo_dis_ei_many_modes": lambda: _eubo(case_dis_ei("many_modes"), 207),
    "ei_cosine": lambda: case_ei_cosine("ei"),
    "em_cosine": lambda: case_ei_cosine("em"),
    "eubo_ei_cosine": lambda: _eubo(case_ei_cosine("ei"), 208),
    "eubo_em_two_modes": lambda: _eubo(case_em_two_modes("score"), 201),
    "eubo_ei_many_modes": lambda: _eubo(case_ei_many_modes(K=100, B=100), 202),
    "eubo_cmcd_gmm": lambda: _eubo(case_cmcd_gmm(), 203),
    "eubo_ddpm_snr": lambda: _eubo(case_ddpm_snr(), 204),
    # 256 particles: the >= 99 % agreement bar of the clamp-mask target then tolerates two particles
    "eubo_cmcd_logreg_sonar": lambda: _eubo(case_cmcd_logreg(166, 60, B=256), 205),
    "eubo_cmcd_logreg_iono": lambda: _eubo(case_cmcd_logreg(280, 33, B=256, ctrl_kind="clipped"), 206),
}


# cases whose TRAINING objective (method='lv') and parameter gradient are pinned by reference-generated fixtures
# (tests/golden/grad_<name>.pt, python -m oracle.make_golden --grads)
GRAD_CASES = ["em_two_modes_score", "ei_many_modes", "ddpm_snr", "ei_phi4_gmm", "pis_many_modes", "dds_many_modes_ito",
              "dds_phi4_ito", "dis_many_modes_ito", "pis_logreg", "dis_many_modes_lerp", "dis_many_modes_langevin",
              "dis_ei_many_modes", "cmcd_gmm", "cmcd_logreg_sonar", "lvtraj_ei_two_modes_gauss"]


def case_mala(target="many_modes"):
    """MALA chains as mcmc_sample starts them (experiments/benchmark_utils.py:268-333): n_chains_per_mode copies of
    every initial point, per-chain adaptive step size; short runs so that the CPU oracle finishes in seconds."""
    if target == "many_modes":
        tgt = many_modes(7, 10)
        x_init, per_mode, h = tgt["loc"].clone(), 4, 0.15
    elif target == "phi4":
        tgt = phi4(24)
        x_init, per_mode, h = torch.stack([torch.ones(24), -torch.ones(24), torch.zeros(24)]), 6, 4e-3
    else:
        tgt = logreg_synthetic(166, 60)
        x_init, per_mode, h = torch.zeros(1, 61), 20, 2e-3
    return {"target": tgt, "x_init": x_init, "n_chains_per_mode": per_mode, "step_size": h, "n_warmup": 15, "n_steps": 45,
            "seed": 300 + len(target)}


def case_rwmh(target="many_modes"):
    """The same chains with mcmc_type='rwmh' (rwmh_step, additions/mcmc.py:258-290); step sizes in the range where the
    random walk accepts about half of its proposals."""
    case = case_mala(target)
    case.update(mcmc_type="rwmh", step_size={"many_modes": 0.25, "phi4": 0.02, "logreg": 0.02}[target], seed=case["seed"] + 50)
    return case


MALA_CASES = {"mala_many_modes": lambda: case_mala("many_modes"), "mala_phi4": lambda: case_mala("phi4"),
              "mala_logreg": lambda: case_mala("logreg"), "rwmh_many_modes": lambda: case_rwmh("many_modes"),
              "rwmh_phi4": lambda: case_rwmh("phi4"), "rwmh_logreg": lambda: case_rwmh("logreg")}


def mala_inputs(case: dict):
    """(y_init [C, d], noise [S, C, d], unif [S, C]) of a MALA case, from the Philox spec of oracle/philox_ref.py."""
    from oracle import philox_ref
    y_init = case["x_init"].repeat_interleave(case["n_chains_per_mode"], dim=0)
    C, d = y_init.shape
    S = case["n_warmup"] + case["n_steps"]
    noise = torch.from_numpy(philox_ref.normals(case["seed"], C, S, d))
    unif = torch.from_numpy(philox_ref.uniforms(case["seed"], C, S))
    return y_init, noise, unif


# training-only cases (not rollout parity cases: their noise covers B * traj_per_sample particles)
TRAIN_ONLY_CASES = {"lvtraj_ei_two_modes_gauss": lambda: _lv_traj(case_ei_two_modes_gauss())}


def grad_case(name: str) -> dict:
    return (CASES.get(name) or TRAIN_ONLY_CASES[name])()


def initial_state(case: dict, dtype=torch.float32):
    """x0 for the case: prior samples for a generative rollout (drawn with the Philox stream 1 of
    oracle.philox_ref so that no torch-generator state is involved), target-shaped samples for EUBO."""
    from oracle import philox_ref
    p = case["problem"]
    d = p["target"]["loc"].shape[1] if p["target"]["kind"] == "gmm" else p["target"]["dim"]
    B = case["B"]
    z = torch.from_numpy(philox_ref.normals(case["seed"], B, 1, d, stream=philox_ref.STREAM_PRIOR)[0]).to(dtype)
    if case.get("eubo"):
        tgt = p["target"]
        if tgt["kind"] == "gmm":  # exact GMM samples: component by inverse-cdf on a Philox uniform-ish draw
            w = tgt["weights"] / tgt["weights"].sum()
            u = torch.from_numpy(philox_ref.normals(case["seed"] + 1, B, 1, 1, stream=philox_ref.STREAM_PRIOR)[0, :, 0])
            u = 0.5 * (1 + torch.erf(u.double() / math.sqrt(2))).float()
            comp = torch.searchsorted(torch.cumsum(w, 0), u.clamp(max=1 - 1e-6))
            return (tgt["loc"][comp] + tgt["scale"][comp] * z).to(dtype)
        return z
    kind = case["prior"][0]
    if kind == "delta":
        return torch.full((B, d), float(case["prior"][1]), dtype=dtype)
    if kind == "iso":
        return (case["prior"][1] + case["prior"][2] * z).to(dtype)
    if kind == "gauss":
        return (p["prior"]["loc"] + p["prior"]["scale"] * z).to(dtype)
    raise ValueError(kind)


def noise_for(case: dict, dtype=torch.float32):
    from oracle import philox_ref
    p = case["problem"]
    d = p["target"]["loc"].shape[1] if p["target"]["kind"] == "gmm" else p["target"]["dim"]
    K = len(p["ts"]) - 1
    return torch.from_numpy(philox_ref.normals(case["seed"], case["B"] * case.get("traj_per_sample", 1), K, d)).to(dtype)
