#!/usr/bin/env python
"""One validation-mode rollout per kernel family (mixture, lattice, logistic-regression, general tensor-core, fp32 SIMT,
MALA) for compute-sanitizer runs:   compute-sanitizer --tool memcheck|racecheck python tools/sanitize_cases.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests.cases import CASES, MALA_CASES, initial_state, mala_inputs, noise_for  # noqa: E402
from tests.product_builders import Built, build_target  # noqa: E402

dev = torch.device("cuda:0")
for name, prec in (("ei_many_modes", "f16x3"), ("ei_close_modes", "f16x3"), ("eubo_ei_many_modes", "f16x3"),
                   ("pis_phi4", "f16x3"), ("cmcd_logreg_sonar", "f16x3"), ("cmcd_gmm", "f16x3"), ("ei_phi4_gmm", "tf32x3"),
                   ("em_two_modes_score", "fp32")):
    case = CASES[name]()
    x0, noise = initial_state(case), noise_for(case)
    built = Built(case, dev, prec)
    if case.get("eubo"):
        rnd = built.compute_eubo(x0, noise)
    else:
        _, rnd, _ = built.simulate(x0, noise)
    torch.cuda.synchronize()
    print(f"{name} [{prec}]: rnd mean {rnd.mean().item():+.5e} finite {bool(torch.isfinite(rnd).all())}", flush=True)
from sde_sampler_lrds_b200.additions.mcmc import mala_chains  # noqa: E402
mc = MALA_CASES["mala_many_modes"]()
y_init, mnoise, unif = mala_inputs(mc)
ys, _ = mala_chains(build_target(mc["target"], dev), y_init.to(dev), mc["step_size"], mc["n_warmup"], mc["n_steps"], noise=mnoise, unif=unif)
torch.cuda.synchronize()
print(f"mala_many_modes: ys mean {ys.mean().item():+.5e}", flush=True)
