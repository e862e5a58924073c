#!/usr/bin/env python
"""Time of the weight-gradient kernel (lrds_mlp_grad) alone on the rows of one BASELINE config-2 training step
(K x B stored states, d = 50), next to the same vector-Jacobian product through torch autograd (cuBLAS):
    python tools/mlp_grad_bench.py [--B 65536] [--K 200] [--json out.json]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=65536)
    ap.add_argument("--K", type=int, default=200)
    ap.add_argument("--d", type=int, default=50)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    from sde_sampler_lrds_b200.models.mlp import FourierMLP
    from sde_sampler_lrds_b200.train import mlp_grad, time_embed_rows
    dev = torch.device("cuda:0")
    S, B, d = args.K, args.B, args.d
    torch.manual_seed(0)
    base = FourierMLP(dim=d, activation=torch.nn.GELU(), num_layers=4).to(dev)
    with torch.no_grad():
        for p in base.parameters():
            if p.requires_grad:
                p.copy_(torch.randn_like(p) * (0.2 if p.ndim == 2 else 0.1))
    xs = torch.randn(S, B, d, device=dev)
    cot = torch.randn(S, B, d, device=dev)
    taus = torch.linspace(0, 1, S, device=dev)
    with torch.no_grad():
        bias1 = time_embed_rows(base.timestep_embed, taus) + base.input_embed.bias

    def timed(fn, n=3):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(n):
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best

    ms_kernel = timed(lambda: mlp_grad(base, bias1, xs, cot, 10.0, cot_bound=6.0))

    def autograd_pass(rows=1 << 20):
        params = [p for p in base.parameters() if p.requires_grad]
        step = max(1, rows // B)
        for k0 in range(0, S, step):
            k1 = min(S, k0 + step)
            emb = base.input_embed(xs[k0:k1]) + time_embed_rows(base.timestep_embed, taus[k0:k1])[:, None, :]
            for layer in base.hidden_layer:
                emb = layer(F.gelu(emb))
            net = base.out_layer(F.gelu(emb)).clip(-10.0, 10.0)
            torch.autograd.grad((cot[k0:k1] * net).sum(), params, allow_unused=True)
    ms_torch = timed(autograd_pass)
    rows = S * B
    flop = rows * 2 * (d * 64 + 2 * 64 * 64 + 64 * d) * 3  # forward, backward-data, weight gradient
    out = {"rows": rows, "d": d, "kernel_ms": ms_kernel, "torch_autograd_ms": ms_torch,
           "rows_per_s": rows / (ms_kernel * 1e-3), "useful_tflops": flop / (ms_kernel * 1e-3) / 1e12,
           "hbm_gb_per_s": rows * d * 8 / (ms_kernel * 1e-3) / 1e9}
    print(json.dumps(out))
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
