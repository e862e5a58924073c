"""GPU: BASELINE.json's full-size shapes.  The oracle cannot run 10^5 particles in seconds, so full-size results are
tied to it through size-independent properties of the path:

  1. particles are independent and the production noise is keyed by the global particle index, hence the first n
     particles of a full-size launch must equal, BIT FOR BIT, a launch of only those n particles;
  2. production mode (in-kernel Philox) must equal, bit for bit, validation mode fed with the same generator's
     normals (lrds_normals), which is the mode the oracle / golden parity tests cover;
  3. that n-particle validation run is compared with the CPU oracle at the north-star tolerance.

Shapes (configs[1..3] of BASELINE.json): ManyModes d=50 EI K=200 B=65536; PhiFour d=100 PIS (EM) and DDS K=256
B=131072; logistic regression (sonar shape, synthetic data) CMCD K=100 B=262144."""
import ctypes as C

import pytest
import torch

from oracle import rollout_oracle as O
from tests import cases as T

pytestmark = pytest.mark.gpu
N_CHECK = 96


def _full_cases():
    return {
        "many_modes_ei": (lambda: T.case_ei_many_modes(K=200, B=65536), "f16x3"),  # the benchmark kernel (bench.py)
        "many_modes_ei_tf32x3": (lambda: T.case_ei_many_modes(K=200, B=65536), "tf32x3"),
        "phi4_pis_f16x3": (lambda: T.case_pis_phi4(K=256, B=131072), "f16x3"),  # rollout_lin_kernel<F16X3, EM>: the kernel the shape table quotes
        "phi4_dds_f16x3": (lambda: _dds256(), "f16x3"),                           # rollout_lin_kernel<F16X3, AXPY>
        "phi4_pis": (lambda: T.case_pis_phi4(K=256, B=131072), "tf32x3"),
        "phi4_dds": (lambda: _dds256(), "bf16"),
        "many_modes_eubo": (lambda: dict(T.case_ei_many_modes(K=200, B=65536), eubo=True), "f16x3"),
        "logreg_cmcd": (lambda: T.case_cmcd_logreg(166, 60, K=100, B=262144), "f16x3"),  # logit / gradient GEMMs on tcgen05
        "logreg_cmcd_fp32": (lambda: T.case_cmcd_logreg(166, 60, K=100, B=32768), "fp32"),
    }


def _dds256():
    case = T.case_dds_phi4(True, B=131072)
    ts = T.cosine_ts(end=6.4, dt=0.025)  # K = 256+ steps: the reference-expressible twin of "256 steps" 
    case["problem"]["ts"] = ts
    return case


@pytest.mark.parametrize("name", list(_full_cases()))
def test_full_size_rollout_is_tied_to_the_oracle(name, device):
    from sde_sampler_lrds_b200 import _native as N
    from tests.product_builders import Built
    make, precision = _full_cases()[name]
    case = make()
    p = case["problem"]
    B = case["B"]
    d = p["target"]["loc"].shape[1] if p["target"]["kind"] == "gmm" else p["target"]["dim"]
    K = len(p["ts"]) - 1
    built = Built(case, device, precision)
    g = torch.Generator().manual_seed(case["seed"])
    if case["prior"][0] == "delta":
        x0 = torch.zeros(B, d)
    else:
        x0 = torch.randn(B, d, generator=g) * float(case["prior"][2]) + float(case["prior"][1])
    seed, off = 0xC0FFEE + case["seed"], 7 * B
    if case.get("eubo"):
        _full_size_eubo(built, case, x0, seed, off, device)
        return
    x_full, rnd_full, _ = built.simulate(x0, None, seed=seed, particle_offset=off)
    assert x_full.shape == (B, d) and rnd_full.shape == (B, 1)
    assert torch.isfinite(x_full).all() and torch.isfinite(rnd_full).all()
    # 1. shard independence, bit for bit (a slice in the middle of the batch, not tile aligned)
    lo = B // 2 + 37
    xs, rs, _ = built.simulate(x0[lo:lo + N_CHECK], None, seed=seed, particle_offset=off + lo)
    assert torch.equal(xs, x_full[lo:lo + N_CHECK])
    if name.startswith("many_modes_ei") and precision == "f16x3":
        # the slice runs the small-batch kernel (four threads per particle, lrds_rollout_mix_small.cuh): identical states,
        # log-weights summed from four partial sums instead of one running sum
        assert ((rs - rnd_full[lo:lo + N_CHECK]).abs() <= 2e-6 * rnd_full[lo:lo + N_CHECK].abs().clamp(min=1.0)).all()
    else:
        assert torch.equal(rs, rnd_full[lo:lo + N_CHECK])
    # 2. production mode == validation mode on the generator's own normals
    noise = torch.empty(K, N_CHECK, d, device=device)
    N.check(N.lib().lrds_normals(C.c_uint64(seed), C.c_uint64(off + lo), 0, K, N_CHECK, d, N.ptr(noise), N.stream_ptr(device)))
    xv, rv, _ = built.simulate(x0[lo:lo + N_CHECK], noise)
    assert torch.equal(xv, xs) and torch.equal(rv, rs)
    # 3. the same increments through the CPU oracle
    xo, ro, _ = O.rollout(p, x0[lo:lo + N_CHECK], noise.cpu(), compute_ito_int=case.get("compute_ito_int", True))
    tol = 1e-2 if precision == "bf16" else 1e-4
    need = 0.99 if (p["target"]["kind"] == "logreg" or precision == "bf16") else 1.0
    ex = ((xv.cpu() - xo).abs() / xo.abs().clamp(min=1.0)).max(dim=1).values
    er = ((rv.cpu() - ro).abs() / ro.abs().clamp(min=1.0)).reshape(-1)
    assert (ex <= tol).float().mean().item() >= need, f"x_T: worst {ex.max().item():.2e}"
    assert (er <= tol).float().mean().item() >= need, f"rnd: worst {er.max().item():.2e}"


def _full_size_eubo(built, case, x0, seed, off, device):
    """The noising rollout (compute_eubo) at full size: shard independence, production == validation mode on the
    generator's normals (two draws per step: the increment enters the update and the cost), oracle parity of the slice."""
    from sde_sampler_lrds_b200 import _native as N
    p, B = case["problem"], case["B"]
    d, K = x0.shape[1], len(p["ts"]) - 1
    rnd_full = built.compute_eubo(x0.clone(), None, seed=seed, particle_offset=off)
    assert rnd_full.shape == (B, 1) and torch.isfinite(rnd_full).all()
    lo = B // 2 + 37
    rs = built.compute_eubo(x0[lo:lo + N_CHECK].clone(), None, seed=seed, particle_offset=off + lo)
    assert torch.equal(rs, rnd_full[lo:lo + N_CHECK])
    noise = torch.empty(K, N_CHECK, d, device=device)
    N.check(N.lib().lrds_normals(C.c_uint64(seed), C.c_uint64(off + lo), 0, K, N_CHECK, d, N.ptr(noise), N.stream_ptr(device)))
    rv = built.compute_eubo(x0[lo:lo + N_CHECK].clone(), noise)
    assert torch.equal(rv, rs)
    ro = O.rollout(p, x0[lo:lo + N_CHECK], noise.cpu(), eubo=True)
    er = ((rv.cpu() - ro).abs() / ro.abs().clamp(min=1.0)).reshape(-1)
    assert er.max().item() <= 1e-4, f"rnd: worst {er.max().item():.2e}"


def test_full_size_training_gradient_is_tied_to_small_batches(device):
    """The LV gradient of a benchmark-size training step (K = 200, B = 65 536: 13.1 M stored states through lrds_mlp_grad
    and lrds_score_cot_sums).  The loss is a variance over the batch, so the gradient is not a sum over particles - but
    the two kernels' outputs ARE sums over rows: with the per-particle weights d loss / d rnd fixed, the full-size pass
    must equal the sum of the passes over the two halves of the batch (size-independent additivity), and one half is
    compared with fp32 torch autograd over the same rows (the pass the fixtures of tests/test_train_gpu.py pin at small
    sizes)."""
    from sde_sampler_lrds_b200 import train as TR
    from tests.product_builders import Built
    K, B, d = 200, 65536, 50
    case = T.case_ei_many_modes(K=K, B=B)
    built = Built(case, device, "f16x3")
    info = built.loss._ctrl(False)
    base = info.base
    g = torch.Generator().manual_seed(5)
    xs = (torch.randn(K, B, d, generator=g) * 3.0).to(device)
    z = torch.randn(K, B, d, generator=g).to(device)
    w = (torch.randn(B, generator=g) * 1e-4).to(device)
    ito = (torch.rand(K, generator=g) * 0.2 + 0.05).to(device)
    taus = torch.linspace(0.01, 0.99, K, device=device)
    with torch.no_grad():
        bias1 = TR.time_embed_rows(base.timestep_embed, taus) + base.input_embed.bias
    bound = 7.0 * float(ito.max()) * float(w.abs().max())
    full, db_full = TR.mlp_grad(base, bias1, xs, z, info.clip_model, ito, w, cot_bound=bound)
    h = B // 2
    parts = [TR.mlp_grad(base, bias1, xs[:, lo:lo + h].contiguous(), z[:, lo:lo + h].contiguous(), info.clip_model, ito,
                         w[lo:lo + h], cot_bound=bound) for lo in (0, h)]
    # (the sums run over 13.1 M random-sign terms in fp32 and in a different order: rounding ~ 6e-8 sqrt(N) ~ 2e-4 of the result)
    for p_, gfull in full.items():
        gsum = parts[0][0][p_] + parts[1][0][p_]
        assert ((gfull - gsum).abs().max() / gfull.abs().max()) < 1e-3, float((gfull - gsum).abs().max() / gfull.abs().max())
    assert ((db_full - (parts[0][1] + parts[1][1])).abs().max() / db_full.abs().max()) < 1e-3
    m_full = TR.score_cot_sums(info.target, xs, z, info.clip_score, ito, w)
    m_half = sum(TR.score_cot_sums(info.target, xs[:, lo:lo + h].contiguous(), z[:, lo:lo + h].contiguous(), info.clip_score,
                                   ito, w[lo:lo + h]) for lo in (0, h))
    assert ((m_full - m_half).abs().max() / m_full.abs().max()) < 1e-3
    # one time slice of the full-size pass against fp32 autograd over the same 65 536 rows
    k = 37
    params = [p_ for p_ in base.parameters() if p_.requires_grad]
    with torch.enable_grad():
        emb = base.input_embed(xs[k]) + TR.time_embed_rows(base.timestep_embed, taus[k:k + 1])
        for layer in base.hidden_layer:
            emb = layer(torch.nn.functional.gelu(emb))
        net = base.out_layer(torch.nn.functional.gelu(emb))
        if info.clip_model is not None:
            net = net.clip(-info.clip_model, info.clip_model)
        ref = dict(zip(params, torch.autograd.grad((z[k] * (ito[k] * w)[:, None] * net).sum(), params, allow_unused=True)))
    one, _ = TR.mlp_grad(base, bias1[k:k + 1], xs[k:k + 1].contiguous(), z[k:k + 1].contiguous(), info.clip_model, ito[k:k + 1],
                         w, cot_bound=bound)
    for p_, gk in one.items():
        assert ((gk - ref[p_]).norm() / ref[p_].norm()) < 1e-3
