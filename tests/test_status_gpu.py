"""GPU: the rollout's status counters (lrds_spec.status, pack.read_status).  The f16x3 precision feeds the tensor cores
fp16 (hi, lo) operands: coordinates beyond 65504 and hidden pre-activations beyond 1023 saturate, and the particle's
result is no longer fp32-grade.  The kernels count such particle threads (and non-finite results) instead of staying
silent; tf32x3 has no magnitude limit and is the precision to rerun with."""
import pytest
import torch

from tests.cases import CASES, initial_state, noise_for

pytestmark = pytest.mark.gpu


def _run(case, device, precision):
    from sde_sampler_lrds_b200 import pack
    from tests.product_builders import Built
    pack.read_status(device)  # clear
    x0, noise = initial_state(case), noise_for(case)
    x, rnd, _ = Built(case, device, precision).simulate(x0, noise)
    return x.cpu(), rnd.cpu(), pack.read_status(device)


@pytest.mark.parametrize("name", ["ei_many_modes", "pis_phi4", "cmcd_logreg_sonar", "cmcd_gmm"])
def test_no_saturation_on_the_parity_cases(name, device):
    _, rnd, st = _run(CASES[name](), device, "f16x3")
    assert torch.isfinite(rnd).all() and st == {"saturated": 0, "nonfinite": 0}


@pytest.mark.parametrize("name", ["ei_many_modes", "pis_phi4", "cmcd_logreg_sonar"])
def test_saturating_activations_are_flagged(name, device):
    """Hidden pre-activations far beyond 1023 (input layer scaled up): f16x3 must flag the particles; tf32x3 runs the
    same problem without a limit and stays at the oracle's result."""
    from oracle import rollout_oracle as O
    case = CASES[name]()
    sd = case["problem"]["ctrl"]["sd"]
    sd["base_model.input_embed.weight"] = sd["base_model.input_embed.weight"] * 4000.0
    sd["base_model.hidden_layer.0.weight"] = sd["base_model.hidden_layer.0.weight"] * 1e-3  # keep the output in range
    B = case["B"]
    _, _, st = _run(case, device, "f16x3")
    assert st["saturated"] >= B // 2, st
    x, rnd, st32 = _run(case, device, "tf32x3")
    assert st32["saturated"] == 0
    x0, noise = initial_state(case), noise_for(case)
    xo, ro, _ = O.rollout(case["problem"], x0, noise, compute_ito_int=case.get("compute_ito_int", True))
    err = ((rnd - ro).abs() / ro.abs().clamp(min=1.0)).reshape(-1)
    need = 0.97 if case["problem"]["target"]["kind"] == "logreg" else 0.99
    assert (err <= 1e-3).float().mean().item() >= need, err.max().item()


def test_saturating_coordinates_and_nonfinite_results_are_flagged(device):
    case = CASES["ei_many_modes"]()
    from sde_sampler_lrds_b200 import pack
    from tests.product_builders import Built
    pack.read_status(device)
    x0, noise = initial_state(case), noise_for(case)
    x0 = x0.clone()
    x0[:7] *= 1e5  # |x| > 65504 in the first seven particles
    x0[9, 3] = float("nan")
    Built(case, device, "f16x3").simulate(x0, noise)
    st = pack.read_status(device)
    assert st["saturated"] >= 7 and st["nonfinite"] >= 1, st
