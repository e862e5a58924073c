"""CPU, world_size 2 over gloo: the only cross-GPU exchange of the path - the all_gather of 8 fp64 estimator
partials per rank and their associative merge (sde_sampler_lrds_b200/estimators.py) - reproduces the
single-process estimators of the oracle (BaseOCLoss.compute_results, losses/oc.py:134-173; ESS,
eval/metrics.py:134-140; evaluate_eubo, additions/hacking.py:24-32) on sharded log-weights."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import rollout_oracle as O
from sde_sampler_lrds_b200 import estimators as E


def host_partials(rnd: torch.Tensor) -> torch.Tensor:
    """The 8 partials of lrds_estimator_partials, stated with torch on the CPU (test double of the kernel)."""
    r = rnd.double().reshape(-1)
    m, m2 = (-r).max(), r.max()
    e = torch.exp(-r - m)
    return torch.stack([m, e.sum(), (e * e).sum(), r.sum(), (r * r).sum(), torch.tensor(float(r.numel()), dtype=torch.float64),
                        m2, torch.exp(r - m2).sum()])


def _worker(rank, world, port, shards, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        merged = E.gather_and_merge(host_partials(shards[rank]), dist.group.WORLD)
        out[rank] = E.metrics_from_partials(merged)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("sizes", [(1000, 1000), (1537, 263)])
def test_two_rank_merge_matches_single_process_estimators(sizes):
    g = torch.Generator().manual_seed(5)
    rnd = (torch.randn(sum(sizes), 1, generator=g) * 3.0 + 40.0).float()
    shards = list(torch.split(rnd, list(sizes)))
    ctx = mp.get_context("spawn")
    with ctx.Manager() as man:
        out = man.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, shards, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        got = [dict(out[r]) for r in range(2)]
    want = O.compute_results(rnd)
    want_f = O.eubo_results(rnd)
    for m in got:  # every rank holds the same global answer
        assert abs(m["elbo"] - want["eval/elbo"]) < 1e-5 * max(1, abs(want["eval/elbo"]))
        assert abs(m["log_norm_const_is"] - want["log_norm_const_is"]) < 1e-4
        assert abs(m["lv_loss"] - want["eval/lv_loss"]) < 1e-4 * max(1, want["eval/lv_loss"])
        assert abs(m["log_norm_const_is_f"] - want_f["eval/log_norm_const_is_f"]) < 1e-4
        assert m["count"] == sum(sizes)
    assert got[0] == got[1]


def test_merge_is_associative_and_ignores_empty_ranks():
    g = torch.Generator().manual_seed(6)
    rnd = torch.randn(4096, generator=g) * 5 - 20
    parts = torch.stack([host_partials(c) for c in rnd.split(1024)])
    whole = host_partials(rnd)
    a = E.merge_partials(parts)
    b = E.merge_partials(torch.stack([E.merge_partials(parts[:2]), E.merge_partials(parts[2:]), torch.zeros(8, dtype=torch.float64)]))
    for i in range(8):
        assert math.isclose(float(a[i]), float(whole[i]), rel_tol=1e-12, abs_tol=1e-12)
        assert math.isclose(float(a[i]), float(b[i]), rel_tol=1e-12, abs_tol=1e-12)


def _lv_worker(rank, world, port, shards, masks, out):
    from sde_sampler_lrds_b200.train import lv_weights
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        loss, w = lv_weights(shards[rank], masks[rank], dist.group.WORLD)
        out[rank] = (loss.item(), w.clone())
    finally:
        dist.destroy_process_group()


def test_two_rank_lv_weights_match_the_single_process_variance_gradient():
    """Training over sharded particles: loss = Var(rnd[mask]) of the GLOBAL batch and d loss / d rnd from one all_reduce
    of three fp64 sums equal autograd through the reference's formula (compute_loss, losses/oc.py:105-131) on the whole
    batch."""
    g = torch.Generator().manual_seed(8)
    sizes = (700, 324)
    rnd = (torch.randn(sum(sizes), 1, generator=g) * 4.0 + 30.0).float()
    rnd[5] = 2e8  # filtered by max_rnd = 1e8
    mask = rnd < 1e8
    leaf = rnd.clone().requires_grad_(True)
    want_loss = leaf[mask].var()
    (want_w,) = torch.autograd.grad(want_loss, leaf)
    ctx = mp.get_context("spawn")
    with ctx.Manager() as man:
        out = man.dict()
        port = _free_port()
        shards, masks = list(torch.split(rnd, list(sizes))), list(torch.split(mask, list(sizes)))
        procs = [ctx.Process(target=_lv_worker, args=(r, 2, port, shards, masks, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        got = [out[r] for r in range(2)]
    assert abs(got[0][0] - want_loss.item()) < 1e-5 * want_loss.item() and got[0][0] == got[1][0]
    w = torch.cat([got[0][1], got[1][1]])
    assert (w - want_w).abs().max() < 1e-6 * want_w.abs().max() and w[5] == 0
