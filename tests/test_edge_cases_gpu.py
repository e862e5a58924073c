"""GPU: ragged and boundary shapes of the specialised tensor-core kernels against the CPU oracle (validation mode,
same increments; 1e-4 relative like tests/test_rollout_parity_gpu.py): single particles, batches that do not fill a
warp / a tile, one-step grids, mixture sizes on every branch of the half-warp mode split (1, 2, 3 and 4 blocks of four
modes), dimensions at and off the 8- and 16-boundaries, lattice lengths at the chunk boundary, data counts at the
16-boundary of the regression GEMMs."""
import pytest
import torch

from oracle import rollout_oracle as O
from tests import cases as T
from tests.cases import initial_state, noise_for

pytestmark = pytest.mark.gpu


def _check(case, device, need=0.97, precision="f16x3"):
    """>= `need` of the particles within 1e-4 of the oracle and EVERY particle within 1e-2: these shapes are chosen for
    their indexing, not their conditioning (a few particles of a 130-particle batch sit on mode boundaries, where the
    fp32 SIMT anchor itself leaves 1e-4), and an indexing bug shows up as a gross error, a NaN or a wrong particle."""
    from tests.product_builders import Built
    x0, noise = initial_state(case), noise_for(case)
    built = Built(case, device, precision)
    if case.get("eubo"):
        rnd = built.compute_eubo(x0, noise).cpu()
        ref = O.rollout(case["problem"], x0, noise, eubo=True)
        pairs = [(rnd, ref)]
    else:
        x, rnd, _ = built.simulate(x0, noise)
        xo, ro, _ = O.rollout(case["problem"], x0, noise, compute_ito_int=case.get("compute_ito_int", True))
        pairs = [(x.cpu(), xo), (rnd.cpu(), ro)]
    for got, want in pairs:
        assert got.shape == want.shape and torch.isfinite(got).all()
        err = (got - want).abs() / want.abs().clamp(min=1.0)
        err = err.max(dim=1).values if err.dim() == 2 else err
        assert err.max().item() <= 1e-2, f"worst {err.max().item():.2e}"
        if err.numel() >= 32:
            assert (err <= 1e-4).float().mean().item() >= need, f"worst {err.max().item():.2e}"


@pytest.mark.parametrize("M", [2, 3, 5, 8, 9, 12, 13, 16])
@pytest.mark.parametrize("d,B,K", [(3, 1, 1), (17, 33, 4), (64, 130, 3)])
def test_mixture_kernel_shapes(M, d, B, K, device):
    case = T.case_ei_many_modes(K=K, B=B, d=d, M=M)
    # the first K steps of a 100-step grid: a K-step grid over the whole horizon is ill-conditioned for every kernel
    # (the fp32 anchor itself leaves 1e-4 there)
    case["problem"]["ts"] = T.uniform_ts(1.0, 100)[:K + 1].clone()
    _check(case, device)
    _check(dict(case, eubo=True, seed=900 + M), device)


@pytest.mark.parametrize("M,d,B,K", [(16, 50, 200, 12), (5, 17, 33, 4), (16, 64, 130, 3), (3, 3, 1, 1), (10, 24, 257, 6)])
def test_small_batch_and_throughput_kernels_agree(M, d, B, K, device, monkeypatch):
    """The benchmark configuration has two kernels: four threads per particle for latency-bound batches
    (lrds_rollout_mix_small.cuh) and one thread per particle for throughput (lrds_rollout_mix.cuh; LRDS_MIX_SMALL=0
    forces it).  Both against the oracle, and against each other: identical states, log-weights to fp32 rounding, in
    validation and in production mode."""
    from tests.product_builders import Built
    case = T.case_ei_many_modes(K=K, B=B, d=d, M=M)
    case["problem"]["ts"] = T.uniform_ts(1.0, 100)[:K + 1].clone()
    x0, noise = initial_state(case), noise_for(case)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("LRDS_MIX_SMALL", mode)
        _check(case, device)
        built = Built(case, device, "f16x3")
        out[mode] = [t.cpu() for t in built.simulate(x0, noise, return_traj=True)] + [t.cpu() for t in built.simulate(x0, None, seed=77)[:2]]
    for i in (0, 2, 3):  # x_T, trajectory (validation mode), x_T (production mode)
        assert torch.equal(out["1"][i], out["0"][i])
    for i in (1, 4):     # log-weights
        assert ((out["1"][i] - out["0"][i]).abs() <= 2e-6 * out["0"][i].abs().clamp(min=1.0)).all()


def test_close_modes_through_both_kernels(device, monkeypatch):
    """The near-tie stress case (logit GEMM hands particles over to the exact quadratic forms) through both kernels."""
    case = T.case_ei_close_modes()
    for mode in ("1", "0"):
        monkeypatch.setenv("LRDS_MIX_SMALL", mode)
        _check(case, device, need=1.0)


@pytest.mark.parametrize("ctrl_kind", ["score", "cancel", "lerp"])
@pytest.mark.parametrize("M,d,B,K,ito", [(2, 3, 1, 1, True), (5, 17, 33, 4, False), (16, 50, 130, 3, True), (9, 64, 200, 2, True)])
def test_dis_mixture_kernel_shapes(ctrl_kind, M, d, B, K, ito, device):
    """DIS over a mixture target on the mixture kernel (configurations 7 and 10): the initial-cost pre-pass, the three
    drift models, with and without the Ito term."""
    case = T.case_dis("many_modes", ito, scale=1.25, ctrl_kind=ctrl_kind)
    p = case["problem"]
    p["target"] = T.many_modes(M, d)
    p["ctrl"] = T.ctrl(d, ctrl_kind, seed=51, out_gain=0.5, gamma=0.02, sde=p["sde"], prior={"loc": 0.0, "scale": 1.25})
    p["ts"] = T.uniform_ts(1.0, 100)[:K + 1].clone()
    case["B"] = B
    _check(case, device)


def test_lattice_target_with_mixture_reference_at_d100(device):
    """RDS with a mixture reference over the d = 100 PhiFour lattice (experiments/sample_phi_four_gmm_mcmc.py): the
    mixture kernel with the wide (224-column) tile layout."""
    case = T.case_ei_phi4_gmm()
    d = 100
    g = torch.Generator().manual_seed(19)
    case["problem"]["target"] = T.phi4(d)
    case["problem"]["ctrl"] = T.ctrl(d, "score", seed=41, out_gain=0.3, gamma=0.004)
    case["problem"]["ref"] = {"kind": "gmm", "means": torch.stack([torch.ones(d), -torch.ones(d), 0.2 * torch.randn(d, generator=g)]),
                              "variances": 0.3 + 0.2 * torch.rand(3, d, generator=g), "weights": torch.tensor([2.0, 2.0, 1.0])}
    case["B"] = 70
    _check(case, device)


@pytest.mark.parametrize("d,B", [(8, 1), (13, 37), (16, 160), (100, 5)])
@pytest.mark.parametrize("which", ["pis", "dds_ito", "dds_noito"])
def test_reference_free_kernel_shapes(which, d, B, device):
    case = T.case_pis_phi4(K=6, B=B) if which == "pis" else T.case_dds_phi4(which == "dds_ito", B=B)
    case["problem"]["target"] = T.phi4(d)
    if which == "pis":
        case["problem"]["ref"] = {"kind": "pis", "loc": torch.zeros(d)}
        case["problem"]["ts"] = T.uniform_ts(5.0, 100)[:7].clone()
    case["problem"]["ctrl"] = T.ctrl(d, "score", seed=31, out_gain=0.3, gamma=0.004)
    if which != "pis":
        case["problem"]["ts"] = T.cosine_ts(6.4, 0.8)
    _check(case, device)


@pytest.mark.parametrize("n,p,B,precision", [(569, 30, 40, "f16x3"), (1000, 24, 33, "f16x3"), (1000, 24, 33, "fp32")])
def test_logistic_regression_at_the_cancer_and_credit_shapes(n, p, B, precision, device):
    """conf/target/cancer.yaml (569 x 30) and credit.yaml (1000 x 24): more data rows than the regression tensor-core
    kernel holds in tensor memory (N <= 288), so these run the general kernels (residuals per particle in shared memory)."""
    case = T.case_cmcd_logreg(n, p, K=4, B=B)
    _check(case, device, need=0.9, precision=precision)
    pis = T.case_pis_logreg(n, p, K=4, B=B)
    pis["problem"]["ts"] = T.uniform_ts(5.0, 100)[:5].clone()
    _check(pis, device, need=0.9, precision=precision)


@pytest.mark.parametrize("n,p,B", [(16, 7, 1), (50, 15, 33), (166, 60, 5), (280, 33, 130)])
def test_logistic_regression_kernel_shapes(n, p, B, device):
    case = T.case_cmcd_logreg(n, p, K=5, B=B)
    _check(case, device, need=0.95)  # clamp-mask flips on top of the conditioning
    _check(dict(case, eubo=True, seed=950 + n), device, need=0.95)
