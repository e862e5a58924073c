#!/usr/bin/env python
"""Time of one LV training objective + gradient (loss(ts, x, ...) and loss.backward()) through the fused rollout, next
to the CPU oracle's autograd through the K-step loop (the reference's way of computing the same gradient), for the
BASELINE config-2 problem at training batch sizes:   python tools/train_bench.py [--json out.json] [--cpu-batch 512]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import rollout_oracle as O  # noqa: E402  (CPU baseline leg only)
from tests import cases as T  # noqa: E402
from tests.product_builders import Built  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--batches", default="512,2048,8192,65536")
    ap.add_argument("--cpu-batch", type=int, default=512)
    ap.add_argument("--K", type=int, default=200)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    rows = []
    for B in [int(b) for b in args.batches.split(",")]:
        case = T.case_ei_many_modes(K=args.K, B=B)
        d = 50
        x0 = torch.randn(B, d, generator=torch.Generator().manual_seed(1)).to(dev)
        built = Built(case, dev, "f16x3")
        params = list(built.ctrl.parameters())

        def step(seed):
            with torch.no_grad():  # an optimiser step changes the weights: the cached plan refreshes its weight images
                params[0].add_(0.0)
            for p in params:
                p.grad = None
            loss, _ = built.train_loss(x0, None, seed=seed)
            loss.backward()
            return loss
        for w in range(2):
            step(w)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(True), torch.cuda.Event(True)) for _ in range(3)]
        for i, (a, b) in enumerate(ev):
            a.record()
            step(10 + i)
            b.record()
        torch.cuda.synchronize()
        ms = min(a.elapsed_time(b) for a, b in ev)
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        built.simulate(x0, None, seed=3)
        b.record()
        torch.cuda.synchronize()
        row = {"shape": f"cfg2 many_modes d=50 M=16 EI K={args.K} LV loss + gradient", "B": B, "gpu_ms": ms,
               "rollout_only_ms": a.elapsed_time(b), "particle_steps_per_s": B * args.K / (ms * 1e-3)}
        rows.append(row)
        print(json.dumps(row), flush=True)
    # CPU: the oracle's autograd through the loop (what the reference's Trainable.step does), all host threads
    B = args.cpu_batch
    case = T.case_ei_many_modes(K=args.K, B=B)
    x0, noise = T.initial_state(case), T.noise_for(case)
    O.lv_loss_and_grads(dict(case["problem"], ts=case["problem"]["ts"][:5]), x0, noise[:4])  # warm-up
    t0 = time.time()
    O.lv_loss_and_grads(case["problem"], x0, noise)
    sec = time.time() - t0
    row = {"shape": f"cfg2 many_modes d=50 M=16 EI K={args.K} LV loss + gradient", "B": B, "cpu_ms": sec * 1e3,
           "cpu_threads": torch.get_num_threads(), "particle_steps_per_s": B * args.K / sec, "kind": "oracle port (CPU autograd)"}
    rows.append(row)
    print(json.dumps(row), flush=True)
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
