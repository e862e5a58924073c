"""lrds_mlp_grad (the weight-gradient kernel, csrc/lrds_mlp_grad.cu) against fp64 autograd through the same network:
FourierMLP.forward (sde_sampler/models/mlp.py:135-143) under the output clip (models/reparam.py:33-43)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _net64(m, taus, xs, bias1=None):
    from sde_sampler_lrds_b200.train import time_embed_rows
    if bias1 is None:
        bias1 = time_embed_rows(m.timestep_embed, taus.double()) + m.input_embed.bias
    emb = F.linear(xs.double(), m.input_embed.weight) + bias1[:, None, :]
    for layer in m.hidden_layer:
        emb = layer(F.gelu(emb))
    return m.out_layer(F.gelu(emb)), bias1


def _reference(m, taus, xs, cot, clip):
    """fp64 autograd: d/d theta of sum <cot, clip(net)>, plus the cotangent of bias1 (m = the fp64 copy)."""
    with torch.no_grad():
        _, bias1 = _net64(m, taus, xs)
    bias1 = bias1.detach().requires_grad_(True)
    net, _ = _net64(m, taus, xs, bias1)
    if clip is not None:
        net = net.clip(-clip, clip)
    (cot.double() * net).sum().backward()
    out = {"input_embed.weight": m.input_embed.weight.grad, "out_layer.weight": m.out_layer.weight.grad,
           "out_layer.bias": m.out_layer.bias.grad}
    for i, layer in enumerate(m.hidden_layer):
        out[f"hidden_layer.{i}.weight"] = layer.weight.grad
        out[f"hidden_layer.{i}.bias"] = layer.bias.grad
    return out, bias1.grad


@pytest.mark.parametrize("d,nh,B,S,clip,weighted", [
    (50, 2, 300, 3, None, False), (16, 2, 128, 2, 0.3, True), (61, 1, 1000, 4, 0.5, False), (2, 0, 77, 1, None, True),
    (50, 2, 4096, 9, 1.0, True), (33, 2, 20000, 40, 0.8, True),
    (50, 2, 1001, 3, 0.9, False),  # odd B: ragged last tiles, and slices that are not 16-byte aligned (no TMA staging)
    (64, 2, 640, 2, None, False), (7, 1, 129, 5, 0.7, True)])
def test_mlp_grad_matches_autograd(d, nh, B, S, clip, weighted):
    from sde_sampler_lrds_b200.models.mlp import FourierMLP
    from sde_sampler_lrds_b200.train import mlp_grad, time_embed_rows
    torch.manual_seed(d * 100 + nh)
    dev = torch.device("cuda:0")
    base = FourierMLP(dim=d, activation=torch.nn.GELU(), num_layers=nh + 2).to(dev)
    with torch.no_grad():
        for p in base.parameters():
            if p.requires_grad:
                p.copy_(torch.randn_like(p) * (1.2 / p.shape[-1] ** 0.5 if p.ndim == 2 else 0.1))
    xs = torch.randn(S, B, d, device=dev) * 1.5
    cot = torch.randn(S, B, d, device=dev)
    taus = torch.linspace(0.0, 1.0, S, device=dev)
    step_w = (torch.rand(S, device=dev) + 0.5) * 0.1 if weighted else None
    row_w = torch.randn(B, device=dev) * 1e-4 if weighted else None
    import copy
    m64 = copy.deepcopy(base).double()
    with torch.no_grad():
        bias1 = time_embed_rows(base.timestep_embed, taus) + base.input_embed.bias
        if clip is not None:  # the clip mask is discontinuous: take the outputs within 1e-3 of the bound out of the sum
            net, _ = _net64(m64, taus, xs)
            assert 0.02 < (net.abs() > clip).double().mean() < 0.98  # the mask matters in this case
            cot = cot * ((net.abs() - clip).abs() > 1e-3).float()
    grads, dbias1 = mlp_grad(base, bias1, xs, cot, clip, step_w=step_w, row_w=row_w)
    torch.cuda.synchronize()
    eff = cot
    if weighted:
        eff = cot * step_w[:, None, None] * row_w[None, :, None]
    ref, ref_db = _reference(m64, taus, xs, eff, clip)
    names = {p: n for n, p in base.named_parameters()}
    assert len(grads) == len(ref)
    # the weight-gradient operands are single fp16 roundings (unbiased, ~3e-4 each); with the random-sign cotangents of
    # this test the sums themselves grow like sqrt(rows), so the relative error stays near 3e-4 for every size
    tol = 1e-3 if S * B >= 500 else 4e-3
    for p, g in grads.items():
        r = ref[names[p]]
        err = (g.double() - r).norm() / r.norm().clamp_min(1e-30)
        assert err < tol, (names[p], float(err))
    err = (dbias1.double() - ref_db).norm() / ref_db.norm()
    assert err < tol, ("dbias1", float(err))
    # fixed summation order: bit-identical on a second call
    grads2, dbias2 = mlp_grad(base, bias1, xs, cot, clip, step_w=step_w, row_w=row_w)
    assert all(torch.equal(grads[p], grads2[p]) for p in grads) and torch.equal(dbias1, dbias2)


def test_mlp_grad_unsupported_shapes_raise():
    from sde_sampler_lrds_b200 import _native as N
    assert N.lib().lrds_mlp_grad_floats(65, 2) < 0 and N.lib().lrds_mlp_grad_floats(50, 3) < 0
    assert N.lib().lrds_mlp_grad_floats(50, 2) == 50 * 64 + 2 * 64 * 64 + 2 * 64 + 64 * 56 + 56


@pytest.mark.parametrize("name,S,B,clip", [("ei_many_modes", 5, 1000, 3.0), ("ei_many_modes", 3, 77, None),
                                           ("pis_phi4", 4, 300, 0.5), ("cmcd_logreg_sonar", 2, 200, 10.0)])
def test_score_cot_sums_matches_elementwise(name, S, B, clip):
    """lrds_score_cot_sums (scores evaluated and reduced in the kernel) against Distribution.score + torch sums."""
    from sde_sampler_lrds_b200.train import score_cot_sums
    from tests.cases import CASES
    from tests.product_builders import build_target
    dev = torch.device("cuda:0")
    target = build_target(CASES[name]()["problem"]["target"], dev)
    d = target.dim
    torch.manual_seed(3)
    xs = torch.randn(S, B, d, device=dev) * 0.7
    cot = torch.randn(S, B, d, device=dev)
    step_w = torch.rand(S, device=dev) + 0.5
    row_w = torch.randn(B, device=dev)
    got = score_cot_sums(target, xs, cot, clip, step_w, row_w)
    sc = target.score(xs.reshape(-1, d)).reshape(S, B, d).double()
    if clip is not None:
        sc = sc.clip(-clip, clip)
    want = (sc * cot.double() * step_w.double()[:, None, None] * row_w.double()[None, :, None]).sum(1)
    assert ((got.double() - want).abs().max() / want.abs().max()) < 1e-5
    assert torch.equal(got, score_cot_sums(target, xs, cot, clip, step_w, row_w))  # fixed summation order
