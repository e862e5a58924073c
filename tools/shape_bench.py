#!/usr/bin/env python
"""Kernel time of the BASELINE.json shapes other than the bench workload (production mode, in-kernel noise), per
precision:   python tools/shape_bench.py [--precisions f16x3,tf32x3,bf16,fp32] [--json out.json]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from tests import cases as T  # noqa: E402
from tests.product_builders import Built  # noqa: E402


def dds256():
    case = T.case_dds_phi4(True, B=131072)
    case["problem"]["ts"] = T.cosine_ts(end=6.4, dt=0.025)
    return case


def cmcd_many_modes():
    case = T.case_cmcd_gmm()
    d = 50
    case["problem"]["target"] = T.many_modes(16, d)
    case["problem"]["ctrl"] = T.ctrl(d, "score", seed=18, out_gain=0.5, gamma=0.02)
    case["problem"]["ts"] = T.uniform_ts(1.0, 200)
    case["problem"]["prior"] = {"loc": torch.zeros(d), "scale": 5.0 * torch.ones(d), "isotropic": True}
    case["prior"] = ("iso", 0.0, 5.0)
    case["B"] = 65536
    return case


def pis_many_modes():
    case = T.case_pis_many_modes()
    d = 50
    case["problem"]["target"] = T.many_modes(16, d)
    case["problem"]["ctrl"] = T.ctrl(d, "score", seed=27, out_gain=0.5, gamma=0.02)
    case["problem"]["ref"] = {"kind": "pis", "loc": torch.zeros(d)}
    case["problem"]["ts"] = T.uniform_ts(5.0, 200)
    case["B"] = 65536
    return case


def dis(target, ctrl_kind, B, K=200):
    case = T.case_dis(target, True, ctrl_kind=ctrl_kind)
    p = case["problem"]
    if target == "many_modes":
        d = 50
        p["target"] = T.many_modes(16, d)
        p["ctrl"] = T.ctrl(d, ctrl_kind, seed=31, out_gain=0.5, gamma=0.02, sde=p["sde"], prior={"loc": 0.0, "scale": 1.0})
    elif target == "phi4":
        d = 100
        p["target"] = T.phi4(d)
        p["ctrl"] = T.ctrl(d, ctrl_kind, seed=32, out_gain=0.3, gamma=0.004, sde=p["sde"], prior={"loc": 0.0, "scale": 1.0})
    p["ts"] = T.uniform_ts(1.0, K)
    case["B"] = B
    return case


def fitted_ref(B):
    """Config 2 with a reference whose components carry their OWN variances and slightly displaced means - what
    fit_gmm (sklearn EM on MCMC samples, experiments/benchmark_utils.py:336-370) returns: the reference's
    responsibilities then take the exact quadratic forms (the logit GEMM serves shared-variance mixtures)."""
    case = T.case_ei_many_modes(K=200, B=B)
    g = torch.Generator().manual_seed(9)
    ref = case["problem"]["ref"]
    ref["variances"] = ref["variances"] * (1.0 + 0.3 * (torch.rand(ref["variances"].shape, generator=g) - 0.5))
    ref["means"] = ref["means"] + 0.05 * torch.randn(ref["means"].shape, generator=g)
    return case


SHAPES = {
    "cfg1 two_modes d=2 EM K=100 B=2048": lambda: dict(T.case_em_two_modes("score"), B=2048),
    "cfg2 many_modes d=50 M=16 EI K=200 B=65536": lambda: T.case_ei_many_modes(K=200, B=65536),
    # the reference's real operating points: evaluation batch 8192, training batches 512 .. 2048
    "small cfg2 many_modes d=50 M=16 EI K=200 B=8192": lambda: T.case_ei_many_modes(K=200, B=8192),
    "small cfg2 many_modes d=50 M=16 EI K=200 B=2048": lambda: T.case_ei_many_modes(K=200, B=2048),
    "small cfg2 many_modes d=50 M=16 EI K=200 B=512": lambda: T.case_ei_many_modes(K=200, B=512),
    "fitted-ref cfg2 many_modes d=50 M=16 EI K=200 B=65536 (per-mode reference variances)": lambda: fitted_ref(65536),
    "fitted-ref cfg2 many_modes d=50 M=16 EI K=200 B=8192 (per-mode reference variances)": lambda: fitted_ref(8192),
    "cfg3 phi4 d=100 PIS K=256 B=131072": lambda: T.case_pis_phi4(K=256, B=131072),
    "cfg3 phi4 d=100 DDS K=256 B=131072": dds256,
    "cfg4 logreg sonar d=61 CMCD K=100 B=262144": lambda: T.case_cmcd_logreg(166, 60, K=100, B=262144),
    "cfg4 logreg iono d=34 CMCD K=100 B=262144": lambda: T.case_cmcd_logreg(280, 33, K=100, B=262144),
    "logreg sonar d=61 PIS K=100 B=262144": lambda: T.case_pis_logreg(166, 60, K=100, B=262144),
    "many_modes d=50 M=16 CMCD K=200 B=65536": lambda: cmcd_many_modes(),
    "many_modes d=50 M=16 PIS K=200 B=65536": lambda: pis_many_modes(),
    # DIS (Bridge + TimeReversalLoss) with its three drift models
    "many_modes d=50 M=16 DIS ScoreCtrl K=200 B=65536": lambda: dis("many_modes", "score", 65536),
    "many_modes d=50 M=16 DIS LerpCtrl K=200 B=65536": lambda: dis("many_modes", "lerp", 65536),
    "many_modes d=50 M=16 DIS CancelDriftCtrl K=200 B=65536": lambda: dis("many_modes", "cancel", 65536),
    "phi4 d=100 DIS ScoreCtrl K=256 B=131072": lambda: dis("phi4", "score", 131072, K=256),
    "logreg sonar d=61 DIS ScoreCtrl K=100 B=262144": lambda: dis("logreg", "score", 262144, K=100),
    # compute_eubo (the noising rollout behind evaluate_eubo) of the same shapes
    "eubo cfg2 many_modes d=50 M=16 EI K=200 B=65536": lambda: dict(T.case_ei_many_modes(K=200, B=65536), eubo=True),
    "eubo cfg4 logreg sonar d=61 CMCD K=100 B=262144": lambda: dict(T.case_cmcd_logreg(166, 60, K=100, B=262144), eubo=True),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precisions", default="f16x3,tf32x3,bf16,fp32")
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    rows = []
    for name, make in SHAPES.items():
        if args.only and args.only not in name:
            continue
        case = make()
        p = case["problem"]
        B = case["B"]
        d = p["target"]["loc"].shape[1] if p["target"]["kind"] == "gmm" else p["target"]["dim"]
        K = len(p["ts"]) - 1
        g = torch.Generator().manual_seed(1)
        x0 = (torch.zeros(B, d) if case["prior"][0] == "delta" else torch.randn(B, d, generator=g)).to(dev)
        if case.get("eubo") and p["target"]["kind"] == "gmm":  # compute_eubo starts at TARGET samples (hacking.py:14-19)
            t = p["target"]
            idx = torch.multinomial(t["weights"] / t["weights"].sum(), B, replacement=True, generator=g)
            x0 = (t["loc"][idx] + t["scale"][idx] * torch.randn(B, d, generator=g)).to(dev)
        for prec in args.precisions.split(","):
            row = {"shape": name, "precision": prec, "B": B, "K": K, "d": d}
            try:
                built = Built(case, dev, prec)
                run = (lambda seed: built.compute_eubo(x0, None, seed=seed)) if case.get("eubo") else \
                    (lambda seed: built.simulate(x0, None, seed=seed))
                for w in range(2):
                    run(w)
                torch.cuda.synchronize()
                ev = [(torch.cuda.Event(True), torch.cuda.Event(True)) for _ in range(3)]
                for i, (a, b) in enumerate(ev):
                    a.record()
                    run(10 + i)
                    b.record()
                torch.cuda.synchronize()
                ms = min(a.elapsed_time(b) for a, b in ev)
                row.update(ms=ms, particle_steps_per_s=B * K / (ms * 1e-3))
            except Exception as e:
                row["error"] = f"{type(e).__name__}: {e}"[:200]
            rows.append(row)
            print(json.dumps(row), flush=True)
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
