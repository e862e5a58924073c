"""Training objective through the fused rollout: ``loss(ts, x, ...)`` of the linear rollout losses with
method 'lv' / 'lv_traj' (sde_sampler/losses/oc.py:364-394, 1240-1272, 1399-1431 -> compute_loss 105-131), the first
version of SURVEY.md 8f item 1.

With these methods the SDE follows the DETACHED control (generative_and_sde_ctrl, oc.py:83-103), so no gradient flows
along a trajectory: the states x_k are constants and

    d rnd_b / d theta = sum_k  c_k <z_bk, d g(t_k, x_bk) / d theta>          (c_k = the Ito weight of step k)

because the running cost g (sde_ctrl - g / 2) has a vanishing derivative at sde_ctrl = g.  The reference obtains this
by autograd through K sequential steps (~40-150 eager ops each).  Here

  1. the rollout is ONE fused kernel launch (csrc/), fed with the generator's own increments z (lrds_normals) and
     returning the trajectory;
  2. d loss / d rnd comes from the reference's loss formula on the B log-weights;
  3. the parameter gradient is ONE batched evaluation of the control on all K x B stored states with the cotangent
     (d loss / d rnd_b) c_k z_bk: a handful of large library GEMMs (cuBLAS through torch autograd) instead of K small ones.

Step 3 is plain PyTorch on the device - plumbing around the kernel, stated as such; a hand-written weight-gradient
kernel that recomputes the activations from the stored states is the follow-up.  Nothing here
runs on the CPU or imports the oracle.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from . import _native as N
from . import pack


def time_embed_rows(m, t: torch.Tensor) -> torch.Tensor:
    """Differentiable TimeEmbed.forward for a 1-D tensor of times (sde_sampler/models/mlp.py:85-96)."""
    arg = m.timestep_coeff * t.reshape(-1, 1) + m.timestep_phase
    e = torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)
    for layer in m.hidden_layer:
        e = F.gelu(layer(e))
    return m.out_layer(e)


def _clip(v, bound):
    return v if bound is None else v.clip(-bound, bound)


def control_rows(info: pack.CtrlInfo, taus: torch.Tensor, xs: torch.Tensor, score: torch.Tensor | None,
                 coef: torch.Tensor | None = None) -> torch.Tensor:
    """g(t_s, x_sb) for times taus [S] and states xs [S, B, d], differentiable in the parameters of the control:
    FourierMLP.forward (models/mlp.py:135-143) under ClippedCtrl / ScoreCtrl / CancelDriftCtrl / LerpCtrl.forward
    (models/reparam.py:33-43, 112-117, 131-147, 189-199).  ``score`` = the (for LerpCtrl: interpolated) target score at
    xs, a constant because the states carry no gradient; ``coef`` [S, 16] = the time-only table columns of the two DIS
    models (pack.dis_ctrl_rows)."""
    base = info.base
    emb = base.input_embed(xs) + time_embed_rows(base.timestep_embed, taus)[:, None, :]
    for layer in base.hidden_layer:
        emb = layer(F.gelu(emb))
    g = _clip(base.out_layer(F.gelu(emb)), info.clip_model)
    if info.kind == N.CTRL_CLIPPED:
        return g
    sc = info.scale_score * _clip(score, info.clip_score)
    if info.score_model is not None:
        sc = sc * _clip(time_embed_rows(info.score_model, taus), info.clip_model)[:, None, :]
    if info.kind == N.CTRL_CANCEL_DRIFT:  # ctrl + drift / diff + 0.5 diff score
        return g + coef[:, N.STEP_CX, None, None] * xs + coef[:, N.STEP_GSCALE, None, None] * sc
    if info.kind == N.CTRL_LERP:          # ctrl + diff score
        return g + coef[:, N.STEP_GSCALE, None, None] * sc
    return g + sc


class _InjectGrads(torch.autograd.Function):
    """loss value whose backward hands the precomputed parameter gradients to autograd."""

    @staticmethod
    def forward(ctx, value, grads, *params):
        ctx.grads = grads
        return value.clone()

    @staticmethod
    def backward(ctx, grad_out):
        return (None, None, *[None if g is None else grad_out * g for g in ctx.grads])


def normals(seed: int, particle_offset: int, K: int, B: int, d: int, device) -> torch.Tensor:
    """The increments z [K, B, d] the production-mode rollout would draw in-kernel (lrds_normals, stream 0)."""
    z = torch.empty(K, B, d, device=device, dtype=torch.float32)
    with torch.cuda.device(device):
        N.check(N.lib().lrds_normals(C.c_uint64(seed & (2 ** 64 - 1)), C.c_uint64(particle_offset), 0, K, B, d, N.ptr(z),
                                     N.stream_ptr(device)))
    return z


def lv_weights(rnd: torch.Tensor, mask: torch.Tensor, group=None):
    """(loss, d loss / d rnd) of loss = Var(rnd[mask]) (unbiased) over the GLOBAL batch when the particles are sharded
    over the ranks of ``group``: three fp64 sums (sum, count, sum of squares) travel through ONE all_reduce - the
    training-side counterpart of the estimator exchange.  Runs on whatever device ``rnd`` lives on
    (NCCL on the GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    r = rnd.detach().double()
    m = mask.to(r.dtype)
    stats = torch.stack([(r * m).sum(), m.sum(), (r * r * m).sum()])
    if group is not None and dist.is_available() and dist.is_initialized():
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    s, n, s2 = stats
    mean = s / n
    var = (s2 - n * mean * mean) / (n - 1.0)
    w = 2.0 * (r - mean) / (n - 1.0) * m
    return var.to(rnd.dtype), w.to(rnd.dtype)


def lv_objective(loss_obj, plan: pack.Plan, info: pack.CtrlInfo, x: torch.Tensor, seed: int, noise=None,
                 particle_offset: int = 0, max_rows: int = 1 << 20, group=None):
    """(loss, metrics) like ``BaseOCLoss.__call__``: ``loss`` is a scalar whose ``backward()`` leaves the LV gradient in
    the ``.grad`` of the control's parameters.  With ``group`` (a torch.distributed group; method 'lv') the batch is the
    union of the ranks' shards (``particle_offset`` = this rank's first global particle index): the variance and its
    weights are global (lv_weights) and the parameter gradients are summed over the ranks, so every rank returns the
    loss and gradient of the whole batch."""
    dev = x.device
    B, d = x.shape
    K = plan.noise_steps
    z = normals(seed, particle_offset, K, B, d, dev) if noise is None else noise.detach().to(dev, torch.float32).contiguous()
    with torch.no_grad():
        x_T, rnd, xs = pack.run_rollout(plan, x, z, seed, particle_offset, True)
    if group is not None:
        if loss_obj.method != "lv":
            raise NotImplementedError("sharded training is built for method 'lv'")
        mask = loss_obj.filter(rnd, samples=x_T)
        loss_obj.n_filtered += (mask.numel() - mask.sum()).item()
        value, w = lv_weights(rnd, mask, group)
        metrics = {"train/n_filtered_cumulative": loss_obj.n_filtered}
    else:
        rnd_leaf = rnd.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            value, metrics = loss_obj.compute_loss(rnd_leaf, samples=x_T)
            (w,) = torch.autograd.grad(value, rnd_leaf)  # d loss / d rnd_b, zero for filtered particles
    params = [p for p in loss_obj.generative_ctrl.parameters() if p.requires_grad]
    grads: list = [None] * len(params)
    taus = plan.taus.to(dev)
    ito_w = plan.ito_w.to(dev)
    coef = torch.zeros(K, N.STEP_BIAS1)
    pack.dis_ctrl_rows(info, plan.taus, coef)
    coef = coef.to(dev)
    step_rows = max(1, max_rows // B)
    for k0 in range(0, K, step_rows):
        k1 = min(K, k0 + step_rows)
        xs_c = xs[k0:k1]
        score = None
        if info.kind != N.CTRL_CLIPPED:
            score = info.target.score(xs_c.reshape(-1, d)).reshape(k1 - k0, B, d)
            if info.kind == N.CTRL_LERP:  # clipped_interpolated_score, models/reparam.py:170-183
                prior = info.prior.score(xs_c.reshape(-1, d)).reshape(k1 - k0, B, d)
                score = torch.lerp(prior, score, coef[k0:k1, N.STEP_LERP, None, None])
        cot = (ito_w[k0:k1, None, None] * w[None, :, :]) * z[k0:k1]
        with torch.enable_grad():
            surrogate = (cot * control_rows(info, taus[k0:k1], xs_c, score, coef[k0:k1])).sum()
        for i, g in enumerate(torch.autograd.grad(surrogate, params, allow_unused=True)):
            if g is not None:
                grads[i] = g if grads[i] is None else grads[i] + g
    if group is not None:  # one all_reduce over the flattened gradients
        import torch.distributed as dist
        flat = torch.cat([(g if g is not None else torch.zeros_like(p)).reshape(-1) for g, p in zip(grads, params)])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        grads = [c.reshape(p.shape) for c, p in zip(flat.split([p.numel() for p in params]), params)]
    return _InjectGrads.apply(value.detach(), grads, *params), metrics


def _accumulate(grads, params, surrogate):
    for i, g in enumerate(torch.autograd.grad(surrogate, params, allow_unused=True)):
        if g is not None:
            grads[i] = g if grads[i] is None else grads[i] + g


def cmcd_lv_objective(loss_obj, plan: pack.Plan, info: pack.CtrlInfo, x: torch.Tensor, seed: int, noise=None,
                      particle_offset: int = 0, max_rows: int = 1 << 20):
    """The same for ControlledLangevinSDELoss (oc.py:666-755, 830-860) with method 'lv'.  Every trajectory point
    (ts[j], xs[j]) enters the log-weight twice - as the start of step j (control u_s) and as the end of step j - 1
    (control u_t) - through cost_k = (drift_s + drift_t) / sigma + u_s - u_t:

        d rnd_b / d theta = sum_j < [j < K] db_j  -  [j >= 1] (cost_{j-1} dt_{j-1} + db_{j-1}),  d g(ts[j], xs[j]) / d theta >

    (the term cost (sde_ctrl - u_s) dt of oc.py:739 removes the cost dt contribution of the start point).  One batched
    pass over the K + 1 stored states evaluates the controls and the tempered drifts for the costs, a second one carries
    the cotangents back."""
    dev = x.device
    B, d = x.shape
    K = plan.noise_steps
    z = normals(seed, particle_offset, K, B, d, dev) if noise is None else noise.detach().to(dev, torch.float32).contiguous()
    with torch.no_grad():
        x_T, rnd, xs = pack.run_rollout(plan, x, z, seed, particle_offset, True)
    rnd_leaf = rnd.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        value, metrics = loss_obj.compute_loss(rnd_leaf, samples=x_T)
        (w,) = torch.autograd.grad(value, rnd_leaf)
    params = [p for p in loss_obj.generative_ctrl.parameters() if p.requires_grad]
    grads: list = [None] * len(params)
    sde = loss_obj.sde
    ts = plan.taus.to(dev)                      # [K + 1]
    dt = (plan.taus[1:] - plan.taus[:-1]).to(dev)
    sig = sde.diff_coeff.to(dev)
    frac = (plan.taus / sde.terminal_t.detach().cpu()).to(dev)
    target, prior = sde.target_score.__self__, sde.prior_score.__self__
    rows = max(1, max_rows // B)
    gd, drift, tscore = [], [], []
    with torch.no_grad():  # pass 1: controls and tempered drifts (eq/sdes.py:101-110) at every stored state
        for j0 in range(0, K + 1, rows):
            j1 = min(K + 1, j0 + rows)
            flat = xs[j0:j1].reshape(-1, d)
            tsc = target.score(flat).reshape(j1 - j0, B, d)
            dr = tsc * frac[j0:j1, None, None] + prior.score(flat).reshape(j1 - j0, B, d) * (1.0 - frac[j0:j1, None, None])
            dr = _clip(dr * (0.5 * sig ** 2), sde.clip_score)
            gd.append(control_rows(info, ts[j0:j1], xs[j0:j1], tsc if info.kind != N.CTRL_CLIPPED else None))
            drift.append(dr)
            tscore.append(tsc)
        gd, drift, tscore = torch.cat(gd), torch.cat(drift), torch.cat(tscore)
        db = dt.sqrt()[:, None, None] * z                                   # [K, B, d]
        cost = (drift[:-1] + drift[1:]) / sig + gd[:-1] - gd[1:]             # [K, B, d]
        cot = torch.zeros_like(xs)
        cot[:-1] += db
        cot[1:] -= cost * dt[:, None, None] + db
        cot *= w[None, :, :]
    for j0 in range(0, K + 1, rows):  # pass 2: cotangents through the control
        j1 = min(K + 1, j0 + rows)
        with torch.enable_grad():
            g = control_rows(info, ts[j0:j1], xs[j0:j1], tscore[j0:j1] if info.kind != N.CTRL_CLIPPED else None)
            _accumulate(grads, params, (cot[j0:j1] * g).sum())
    return _InjectGrads.apply(value.detach(), grads, *params), metrics
