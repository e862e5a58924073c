import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_mlp_grad_gpu import _reference
from sde_sampler_lrds_b200.models.mlp import FourierMLP
from sde_sampler_lrds_b200.train import mlp_grad, time_embed_rows

def run(d, nh, B, S, clip, weighted):
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
    with torch.no_grad():
        bias1 = time_embed_rows(base.timestep_embed, taus) + base.input_embed.bias
    grads, dbias1 = mlp_grad(base, bias1, xs, cot, clip, step_w=step_w, row_w=row_w)
    eff = cot * step_w[:, None, None] * row_w[None, :, None] if weighted else cot
    import copy
    m64 = copy.deepcopy(base).double()
    ref, ref_db = _reference(m64, taus, xs, eff, clip)
    names = {p: n for n, p in base.named_parameters()}
    out = {names[p]: float((g.double() - ref[names[p]]).norm() / ref[names[p]].norm()) for p, g in grads.items()}
    out["dbias1"] = float((dbias1.double() - ref_db).norm() / ref_db.norm())
    print((d, nh, B, S, clip, weighted), {k: f"{v:.1e}" for k, v in out.items()}, flush=True)

for cfg in [(50, 2, 128, 1, None, False), (50, 2, 4096, 9, None, False), (50, 2, 4096, 9, 1.0, False), (50, 2, 4096, 9, 1.0, True),
            (50, 2, 4096, 9, None, True), (50, 2, 128 * 148 * 2, 1, None, False), (33, 2, 20000, 40, 0.8, True), (50, 2, 65536, 20, 10.0, False)]:
    run(*cfg)
