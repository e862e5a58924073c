#!/usr/bin/env python
"""Where one LV training step (train.lv_objective) of the BASELINE config-2 problem spends its time:
    python tools/train_phases.py [--B 65536] [--K 200]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from tests import cases as T  # noqa: E402
from tests.product_builders import Built  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=65536)
    ap.add_argument("--K", type=int, default=200)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    from sde_sampler_lrds_b200 import train as TR
    dev = torch.device("cuda:0")
    case = T.case_ei_many_modes(K=args.K, B=args.B)
    x0 = torch.randn(args.B, 50, generator=torch.Generator().manual_seed(1)).to(dev)
    built = Built(case, dev, "f16x3")
    for w in range(2):
        loss, _ = built.train_loss(x0, None, seed=w)
        loss.backward()
    torch.cuda.synchronize()
    marks = []
    orig = {n: getattr(TR, n) for n in ("normals", "mlp_grad", "control_param_grads")}
    orig_roll = TR.pack.run_rollout

    def wrap(name, fn):
        def inner(*a, **k):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            marks.append((name, e0, e1))
            return out
        return inner
    for n, f in orig.items():
        setattr(TR, n, wrap(n, f))
    TR.pack.run_rollout = wrap("run_rollout(noise, traj)", orig_roll)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    loss, _ = built.train_loss(x0, None, seed=7)
    loss.backward()
    e1.record()
    torch.cuda.synchronize()
    out = {"B": args.B, "K": args.K, "step_ms": e0.elapsed_time(e1)}
    for name, a, b in marks:
        out[name + "_ms"] = a.elapsed_time(b)
    print(json.dumps(out))
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
