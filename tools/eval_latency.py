#!/usr/bin/env python
"""Wall-clock time of one evaluation through the solver API (make_model(...).compute_results(): prior.sample ->
loss.eval -> estimators -> metrics; the reference's solver/oc.py:148-190) next to the kernel time of its rollout, for
the BASELINE config-2 problem at the reference's evaluation batch:   python tools/eval_latency.py [--B 8192] [--K 200]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8192)
    ap.add_argument("--K", type=int, default=200)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    from sde_sampler_lrds_b200 import benchmark_utils as BU
    from tests.test_solver_api_gpu import _gmm_ref, _randomise_last_layers
    d, M = 50, 16
    details = {"sigma": 1.0, **_gmm_ref(d, M)}
    model = BU.make_model(solver_type="vp-ref", ref_type="gmm", loss_type="lv", integrator_type="ei",
                          model_type="target_informed_zero_init", time_type="uniform", solver_details=details,
                          target_details=BU.make_target_details("many_modes", dim=d, n_modes=M),
                          training_details={"train_steps": 4, "train_batch_size": 512, "eval_batch_size": args.B},
                          n_steps=args.K, device="cuda:0", force_vp20=True)
    _randomise_last_layers(model)
    for _ in range(3):
        model.compute_results()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        res = model.compute_results()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    out = {"B": args.B, "K": args.K, "compute_results_wall_ms": wall, "metrics": sorted(res.metrics)[:6]}
    print(json.dumps(out))
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
