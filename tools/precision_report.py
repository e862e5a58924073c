#!/usr/bin/env python
"""Per-case error table of every kernel precision against the CPU oracle on identical Brownian increments
(validation mode).  Run on a GPU box:  python tools/precision_report.py [--json out.json]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import rollout_oracle as O  # noqa: E402
from tests.cases import CASES, initial_state, noise_for  # noqa: E402
from tests.product_builders import Built  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--precisions", default="fp32,tf32x3,tf32,bf16")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    rows = []
    for name, make in CASES.items():
        case = make()
        x0, noise = initial_state(case), noise_for(case)
        if case.get("eubo"):
            ref_rnd = O.rollout(case["problem"], x0, noise, eubo=True)
            ref_x = None
            mref = O.eubo_results(ref_rnd)
        else:
            ref_x, ref_rnd, _ = O.rollout(case["problem"], x0, noise, compute_ito_int=case.get("compute_ito_int", True))
            mref = O.compute_results(ref_rnd)
        for prec in args.precisions.split(","):
            row = {"case": name, "precision": prec}
            try:
                built = Built(case, dev, prec)
                if case.get("eubo"):
                    rnd = built.compute_eubo(x0, noise).cpu()
                    x = None
                    m = O.eubo_results(rnd)
                else:
                    x, rnd, _ = built.simulate(x0, noise)
                    x, rnd = x.cpu(), rnd.cpu()
                    m = O.compute_results(rnd)
                er = (rnd - ref_rnd).abs() / ref_rnd.abs().clamp(min=1.0)
                row["rnd_max_rel"] = er.max().item()
                row["rnd_frac_1e-4"] = (er <= 1e-4).float().mean().item()
                if x is not None:
                    ex = ((x - ref_x).abs() / ref_x.abs().clamp(min=1.0)).max(dim=1).values
                    row["x_max_rel"] = ex.max().item()
                    row["x_frac_1e-4"] = (ex <= 1e-4).float().mean().item()
                for k, v in mref.items():
                    if "log_norm_const" in k:
                        row["dlogZ"] = abs(m[k] - v)
            except Exception as e:  # report, do not hide
                row["error"] = f"{type(e).__name__}: {e}"[:160]
            rows.append(row)
            print(json.dumps(row), flush=True)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
