#!/usr/bin/env python
"""Phase timers of the benchmark kernel (tools only): runs the bench rollout once with the -DLRDS_MIX_TIMING build
(LRDS_B200_LIB must point at it) and prints, per particle warp of CTA 0 and of the middle CTA, the cycles per grid step
spent between the kernel's marks."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sde_sampler_lrds_b200 import _native as N  # noqa: E402
from tests import cases as T  # noqa: E402
from tests.product_builders import Built  # noqa: E402

NAMES = ["full-wait", "x+logits", "Q_tgt", "wait G1", "epi1", "Q_ref", "wait G2", "epi2+", "wait G3+", "store R",
         "chunk wait", "chunk ld", "chunk math", "tail"]
B = int(os.environ.get("B", 65536))
K = 200
case = T.case_ei_many_modes(K=K, B=B)
dev = torch.device("cuda:0")
x0 = torch.randn(B, 50, generator=torch.Generator().manual_seed(1)).to(dev)
built = Built(case, dev, "f16x3")
EUBO = os.environ.get("EUBO", "0") == "1"
for i in range(3):
    if EUBO:
        built.compute_eubo(x0.clone(), None, seed=i)
    else:
        built.simulate(x0, None, seed=i)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 512)()
SMALL = B <= 128 * 148 and os.environ.get("LRDS_MIX_SMALL", "1") != "0" and not EUBO
if SMALL:
    NAMES = ["sync+full", "store_x", "arr LOGIT", "wait LOGIT", "softmax", "arr G_in", "exact Q", "wait GEMMs", "epilogues",
             "wait out", "R+arr chunks", "wait chunks", "chunk math+tail", "-"]
    assert N.lib().lrds_debug_mix_small_timing(buf) == 0
else:
    assert N.lib().lrds_debug_mix_timing(buf) == 0
for cta in range(2):
    print(f"CTA {'0' if cta == 0 else 'mid'}: cycles per grid step")
    print("warp " + " ".join(f"{n:>10s}" for n in NAMES) + "      total")
    for w in range(16):
        row = [buf[(cta * 16 + w) * 16 + i] / K for i in range(14)]
        if sum(row) == 0:
            continue
        print(f"{w:4d} " + " ".join(f"{v:10.0f}" for v in row) + f" {sum(row):10.0f}"
              f"   exact forms: tgt {buf[(cta * 16 + w) * 16 + 14] / K:.3f} ref {buf[(cta * 16 + w) * 16 + 15] / K:.3f} of the steps")
