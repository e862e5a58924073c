#!/usr/bin/env python
"""Time of lrds_score_cot_sums on the rows of one BASELINE config-2 training step (K x B states of ManyModes d = 50):
    python tools/score_cot_bench.py [--B 65536] [--K 200]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests import cases as T  # noqa: E402
from tests.product_builders import build_target  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=65536)
    ap.add_argument("--K", type=int, default=200)
    args = ap.parse_args()
    from sde_sampler_lrds_b200.train import score_cot_sums
    dev = torch.device("cuda:0")
    target = build_target(T.case_ei_many_modes(K=4, B=8)["problem"]["target"], dev)
    xs = torch.randn(args.K, args.B, target.dim, device=dev)
    cot = torch.randn_like(xs)
    sw, rw = torch.rand(args.K, device=dev), torch.randn(args.B, device=dev)
    for _ in range(2):
        score_cot_sums(target, xs, cot, 100.0, sw, rw)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        score_cot_sums(target, xs, cot, 100.0, sw, rw)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(json.dumps({"rows": args.K * args.B, "d": target.dim, "score_cot_sums_ms": best}))


if __name__ == "__main__":
    main()
