#!/usr/bin/env python
"""Phase timers of the weight-gradient kernel (tools only): runs lrds_mlp_grad once with the -DLRDS_MG_TIMING build
(LRDS_B200_LIB must point at it) and prints the cycles per tile between the kernel's marks for threads 0 (the issuing
warp) and 32 of CTA 0."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sde_sampler_lrds_b200 import _native as N  # noqa: E402
from sde_sampler_lrds_b200.models.mlp import FourierMLP  # noqa: E402
from sde_sampler_lrds_b200.train import mlp_grad, time_embed_rows  # noqa: E402

NAMES = ["load x", "wait prev", "store x+hand", "wait F0", "epi0+hand", "wait F1", "epi1+hand", "wait F2", "epi2+hand",
         "wait Fout", "out+hand", "wait B3", "bwd3+hand", "wait B2", "bwd2+hand", "wait B1", "bwd1+hand"]
S, B, d = int(os.environ.get("S", 200)), int(os.environ.get("B", 65536)), 50
dev = torch.device("cuda:0")
torch.manual_seed(0)
base = FourierMLP(dim=d, activation=torch.nn.GELU(), num_layers=4).to(dev)
with torch.no_grad():
    for p in base.parameters():
        if p.requires_grad:
            p.copy_(torch.randn_like(p) * (0.2 if p.ndim == 2 else 0.1))
xs, cot = torch.randn(S, B, d, device=dev), torch.randn(S, B, d, device=dev)
taus = torch.linspace(0, 1, S, device=dev)
with torch.no_grad():
    bias1 = time_embed_rows(base.timestep_embed, taus) + base.input_embed.bias
for _ in range(2):
    mlp_grad(base, bias1, xs, cot, 10.0, cot_bound=6.0)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 48)()
assert N.lib().lrds_debug_mlp_grad_timing(buf) == 0
tiles = S * ((B + 127) // 128) // 148
for t, name in ((0, "thread 0 (issuer warp)"), (1, "thread 32")):
    row = [buf[t * 24 + i] / tiles for i in range(len(NAMES))]
    print(name, f"- cycles per tile (total {sum(row):.0f})")
    for n, v in zip(NAMES, row):
        print(f"  {n:>14s} {v:8.0f}")
