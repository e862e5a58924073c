#!/usr/bin/env python
"""Markdown table from the JSON of tools/precision_report.py:   python tools/precision_md.py <in.json> <round> > profiles/rNN_precision.md"""
import collections
import json
import sys

rows = json.load(open(sys.argv[1]))
R = sys.argv[2] if len(sys.argv) > 2 else "r02"
precs, cases = [], collections.OrderedDict()
for r in rows:
    if r["precision"] not in precs:
        precs.append(r["precision"])
    cases.setdefault(r["case"], {})[r["precision"]] = r
print(f"## {R} precision table: every kernel precision against the CPU oracle on identical Brownian increments (B200)\n")
print(f"`python tools/precision_report.py --precisions {','.join(precs)}` - max relative error of the per-particle log-weights "
      "`rnd` (denominator max(|ref|, 1)) / of the final states `x_T`, and |log Z - oracle|. fp32, tf32x3 and f16x3 are held to "
      "the north-star tolerance (1e-4 relative, 1e-3 absolute on log Z; tests/test_rollout_parity_gpu.py); tf32 and bf16 are "
      "the reduced-precision fast modes reported separately.\n")
print("| case | " + " | ".join(f"{p}: rnd / x_T / dlogZ" for p in precs) + " |")
print("|---|" + "---|" * len(precs))
worst = collections.defaultdict(float)


def fmt(v):
    return "-" if v is None else f"{v:.1e}"


for name, by in cases.items():
    cells = []
    for p in precs:
        r = by.get(p, {})
        if "error" in r:
            cells.append("error: " + r["error"][:40])
            continue
        cells.append(f"{fmt(r.get('rnd_max_rel'))} / {fmt(r.get('x_max_rel'))} / {fmt(r.get('dlogZ'))}")
        if "logreg" not in name:
            worst[p] = max(worst[p], r.get("rnd_max_rel") or 0.0)
    print(f"| {name} | " + " | ".join(cells) + " |")
print("\nworst log-weight error per precision over the cases without the clamp-mask discontinuity of the regression "
      "posterior (those are held to a fraction-of-particles bar, DESIGN.md 2): " + ", ".join(f"{p} {worst[p]:.1e}" for p in precs))
