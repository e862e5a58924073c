"""Bayesian logistic-regression posterior (reference: sde_sampler/distr/logistic_regression.py).

The reference does NOT override ``score`` (lines 91-92 are commented out), so its solvers differentiate
``posterior_log_prob`` through the two clamps (distr/base.py:146-154); the kernel evaluates that masked
gradient in closed form.  Data come from ``data/<data_type>.pkl`` like the reference,
or are passed directly as tensors (``X_train=..., y_train=...``) for synthetic-shape problems."""
from __future__ import annotations

import pickle

import torch

from .. import _native as N
from .base import DATA_DIR, Distribution


def logreg_tc_image(X: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Tensor-core operand of both logistic-regression GEMMs (lrds_logreg.x_tc, include/lrds_b200.h): the matrix
    Img[n][k] = X[n][k] (k < p), 1 (k = p, the intercept), 0 beyond, rows padded to a multiple of 16 data and columns to
    a multiple of 16 features, times the power of two that puts its largest entry into [2^14, 2^15), split into fp16
    hi | lo parts in the K-major no-swizzle tcgen05 layout [k/8][n][k%8]; then 16 bytes whose first float is the
    un-scale, then the labels y as n_pad floats.  The logit GEMM reads it K-major, the gradient GEMM MN-major."""
    n, p = X.shape
    npad, kp = (n + 15) // 16 * 16, (p + 1 + 15) // 16 * 16
    img = torch.zeros(npad, kp, dtype=torch.float64)
    img[:n, :p] = X.double()
    img[:n, p] = 1.0
    k = 14 - int(torch.floor(torch.log2(img.abs().max())))
    vs = (img * 2.0 ** k).float()
    hi = vs.half()
    lo = (vs - hi.float()).half()
    parts = [t.reshape(npad, kp // 8, 8).permute(1, 0, 2).contiguous().reshape(-1).view(torch.uint8) for t in (hi, lo)]
    tail = torch.zeros(4, dtype=torch.float32)
    tail[0] = 2.0 ** -k
    yp = torch.zeros(npad, dtype=torch.float32)
    yp[:n] = y.float()
    return torch.cat(parts + [tail.view(torch.uint8), yp.view(torch.uint8)]).contiguous()


class LogisticRegression(Distribution):
    def __init__(self, dim=None, data_type=None, use_intercept=True, intercept_mean=0.0, intercept_scale=2.5,
                 weight_scale=1.0, threshold=1e-8, X_train=None, y_train=None, data_dir=None, **kwargs):
        if not use_intercept:
            raise NotImplementedError("use_intercept=False is not used by any shipped target config")
        if X_train is None:
            path = (DATA_DIR if data_dir is None else data_dir) / f"{data_type}.pkl"
            with open(path, "rb") as f:
                data = pickle.load(f)
            X_train, y_train = data["X_train"], data["y_train"]
            self.X_test = data["X_test"].float()
            self.y_test = data["y_test"].float().flatten()
        self.X_train = X_train.float()
        self.y_train = y_train.float().flatten()
        super().__init__(dim=self.X_train.shape[-1] + 1, **kwargs)
        self.threshold = 1e-8  # the reference ignores the constructor argument (line 27)
        self.use_intercept = True
        self.register_buffer("weight_scale", torch.tensor(weight_scale), persistent=False)
        self.register_buffer("intercept_mean", torch.tensor(intercept_mean), persistent=False)
        self.register_buffer("intercept_scale", torch.tensor(intercept_scale), persistent=False)

    def _lrds_pack(self, device):
        X = self.X_train.detach().to("cpu", torch.float32)
        n, p = X.shape
        d_pad, n_pad = ((p + 1 + 7) // 8) * 8, ((n + 3) // 4) * 4
        Xp = torch.zeros(n, d_pad)
        Xp[:, :p] = X
        Xt = torch.zeros(p, n_pad)
        Xt[:, :n] = X.T
        y = torch.zeros(n_pad)
        y[:n] = self.y_train.detach().to("cpu", torch.float32)
        keep = (Xp.to(device), Xt.to(device), y.to(device), logreg_tc_image(X, y[:n]).to(device))
        d = N.Distr()
        d.kind = N.DISTR_LOGREG
        L = d.logreg
        L.N, L.p, L.n_pad = n, p, n_pad
        L.X, L.Xt, L.y = keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr()
        L.x_tc = keep[3].data_ptr()
        L.weight_scale = float(self.weight_scale)
        L.intercept_mean = float(self.intercept_mean)
        L.intercept_scale = float(self.intercept_scale)
        L.threshold = float(self.threshold)
        L.eps = float(torch.finfo(torch.float32).eps)
        return d, keep

    def posterior_log_prob(self, params, X=None, y=None):
        if X is not None and X is not self.X_train:
            return LogisticRegression(X_train=X, y_train=y, intercept_mean=float(self.intercept_mean),
                                      intercept_scale=float(self.intercept_scale),
                                      weight_scale=float(self.weight_scale)).unnorm_log_prob(params).squeeze(-1)
        return self.unnorm_log_prob(params).squeeze(-1)

    def _apply(self, fn):
        out = super()._apply(fn)
        self.X_train, self.y_train = fn(self.X_train), fn(self.y_train)
        return out
