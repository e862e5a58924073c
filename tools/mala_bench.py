#!/usr/bin/env python
"""Time of mcmc_sample's MALA loop at the reference's defaults (n_chains_per_mode=4, dataset_length=50000,
n_warmup_steps=512) over the BASELINE config-2 target (ManyModes d=50, 16 modes: 64 chains x 1293 steps): one lrds_mala
launch on the GPU, the oracle port of the Python loop on the host cores:   python tools/mala_bench.py [--json out.json]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import philox_ref, rollout_oracle as O  # noqa: E402  (CPU baseline leg only)
from sde_sampler_lrds_b200.additions.mcmc import mala_chains  # noqa: E402
from tests import cases as T  # noqa: E402
from tests.product_builders import build_target  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--cpu-steps", type=int, default=200)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    rows = []
    for name, tgt, x_init, h in (("many_modes d=50 M=16", T.many_modes(16, 50), None, 0.05),
                                 ("phi4 d=100", T.phi4(100), torch.stack([torch.ones(100), -torch.ones(100)]), 2e-3)):
        x_init = tgt["loc"].clone() if x_init is None else x_init
        y_init = x_init.repeat_interleave(4, dim=0)
        C, d = y_init.shape
        n_steps = 50000 // C
        target = build_target(tgt, dev)
        for w in range(2):
            mala_chains(target, y_init.to(dev), h, 512, n_steps, seed=w)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        ys, hs, acc = mala_chains(target, y_init.to(dev), h, 512, n_steps, seed=7, return_log_acc=True)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        S = 512 + n_steps
        noise = torch.from_numpy(philox_ref.normals(7, C, args.cpu_steps, d))
        unif = torch.from_numpy(philox_ref.uniforms(7, C, args.cpu_steps))
        t0 = time.time()
        O.mala_chains(tgt, y_init, h, 0, args.cpu_steps, noise, unif)
        cpu_ms_per_step = (time.time() - t0) * 1e3 / args.cpu_steps
        row = {"shape": f"MALA {name}, {C} chains x {S} steps (dataset 50000, warm-up 512)", "gpu_ms": ms,
               "gpu_us_per_step": 1e3 * ms / S, "mean_acceptance": torch.exp(acc.clamp(max=0)).mean().item(),
               "cpu_ms_per_step": cpu_ms_per_step, "cpu_ms_extrapolated": cpu_ms_per_step * S,
               "cpu_threads": torch.get_num_threads(), "cpu_kind": f"oracle port, {args.cpu_steps} steps timed"}
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
