"""Fits the single-branch erf used by the kernels GELU (csrc/lrds_device.cuh: gelu_exact): erf(t) = 1 - 2^(t Q(t)) on [0, 4],
minimax in absolute error, and emulates the fp32 evaluation order to report the GELU error.  Degree 6 is what ships."""
import numpy as np
from scipy.special import erf, erfc
from scipy.optimize import least_squares

# erf(t) ~ 1 - 2^(t * Q(t)),  t in [0, T];  beyond T clamp (erf = 1 - tiny)
T = 4.0
ts = np.cos(np.linspace(0, np.pi, 4001)) * 0.5 * T + 0.5 * T
ts = np.sort(np.concatenate([ts, np.linspace(0, 0.2, 400)]))
target = erf(ts)

def model(c, t):
    q = np.zeros_like(t)
    for ck in c[::-1]:
        q = q * t + ck
    return 1.0 - np.exp2(t * q)

for deg in (5, 6, 7, 8):
    # init from log2(erfc)/t
    tt = ts[ts > 1e-3]
    y = np.log2(erfc(tt)) / tt
    w = erfc(tt)
    c0 = np.polynomial.polynomial.polyfit(tt, y, deg, w=w)
    res = least_squares(lambda c: (model(c, ts) - target) * 1e7, c0, method="lm", max_nfev=20000)
    c = res.x
    # crude minimax refinement: iteratively reweighted
    wts = np.ones_like(ts)
    for it in range(60):
        e = model(c, ts) - target
        wts *= (1 + 2.0 * np.abs(e) / np.abs(e).max())
        wts /= wts.mean()
        res = least_squares(lambda cc: (model(cc, ts) - target) * wts * 1e7, c, method="lm", max_nfev=2000)
        c = res.x
    e = model(c, ts) - target
    print(deg, "max abs err", np.abs(e).max())
    print("  coeffs:", ", ".join(f"{v:.9e}" for v in c))

    # fp32 emulation of the GELU pipeline
    def gelu32(v, c=c):
        v = v.astype(np.float32)
        x = (v * np.float32(0.70710678118654752440)).astype(np.float32)
        t = np.minimum(np.abs(x), np.float32(T)).astype(np.float32)
        q = np.full_like(t, np.float32(c[-1]))
        for ck in c[-2::-1]:
            q = (q.astype(np.float64) * t + np.float32(ck)).astype(np.float32)  # fma
        p = (q.astype(np.float64) * t).astype(np.float32)
        e2 = np.exp2(p.astype(np.float64)).astype(np.float32)
        r = (np.float32(1.0) - e2).astype(np.float32)
        r = np.copysign(r, x)
        h = (np.float32(0.5) * v).astype(np.float32)
        return (h.astype(np.float64) * r + h).astype(np.float32)
    vs = np.linspace(-8, 8, 400001)
    ref = 0.5 * vs * (1 + erf(vs / np.sqrt(2)))
    g = gelu32(vs)
    err = np.abs(g - ref)
    # torch-like fp32 reference error for comparison
    x32 = (vs.astype(np.float32) * np.float32(0.70710678118654752440))
    ref32 = (np.float32(0.5) * vs.astype(np.float32) * (np.float32(1) + erf(x32.astype(np.float64)).astype(np.float32))).astype(np.float32)
    print("  gelu fp32 max abs err", err.max(), "at", vs[err.argmax()], " rel-to-max(1,|v|):", (err / np.maximum(1, np.abs(vs))).max(),
          " (plain fp32 erf formula:", np.abs(ref32 - ref).max(), ")")
