"""Host-buffer front end of the fused rollout: ``HostRolloutStream`` takes the initial particles from pinned host
memory and returns the final particles, their log-weights and the (merged) estimator record in pinned host memory.

A caller that keeps its particles on the host (the reference's evaluation loop does: ``prior.sample`` -> ``simulate`` ->
``.cpu()`` metrics, solver/oc.py:148-190) would otherwise pay H2D copy + kernel + D2H copy back to back for every
rollout.  Here every submitted rollout owns one of ``depth`` slots; its three stages run on three CUDA streams

    copy-in stream   x0 (pinned host) -> slot's device buffer
    compute stream   lrds_rollout, lrds_estimator_partials, and with a process group the all_gather of the 8 fp64
                     partials per rank + lrds_estimator_merge (the ONLY exchange of the path) - all on the device
    copy-out stream  x_T, rnd and the 8 merged doubles -> the slot's pinned host buffers

chained by events, so the copies of rollout i + 1 / i - 1 overlap the kernel of rollout i and nothing synchronises the
host except ``wait(ticket)`` on the one event of the rollout whose results the caller wants.
"""
from __future__ import annotations

import torch

from . import _native as N
from .estimators import gather_and_merge


class HostRolloutStream:
    """``simulate(x_dev, seed, particle_offset) -> (x_T, rnd)`` is the device-resident rollout (e.g. a closure over
    ``loss.simulate``); ``B`` x ``d`` the particle block every ``submit`` carries."""

    def __init__(self, simulate, B: int, d: int, device, group=None, depth: int = 2):
        self.simulate, self.B, self.d, self.device, self.group, self.depth = simulate, B, d, torch.device(device), group, depth
        if self.device.type != "cuda":
            raise N.LrdsError("HostRolloutStream drives the CUDA rollout (no CPU fallback)")
        self.s_in, self.s_out = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        self.x_dev = [torch.empty(B, d, device=self.device) for _ in range(depth)]
        self.x_host = [torch.empty(B, d).pin_memory() for _ in range(depth)]
        self.rnd_host = [torch.empty(B, 1).pin_memory() for _ in range(depth)]
        self.part_host = [torch.empty(8, dtype=torch.float64).pin_memory() for _ in range(depth)]
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_comp = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self._live = [None] * depth    # device results of the slot, kept until its copy-out has been waited for
        self._busy = [False] * depth
        self._n = 0
        blocks = N.lib().lrds_estimator_blocks(B)
        self._scratch = [torch.empty(8 * (blocks + 1), device=self.device, dtype=torch.float64) for _ in range(depth)]
        self.h2d_bytes, self.d2h_bytes = B * d * 4, B * d * 4 + B * 4 + 64

    def submit(self, x0_host: torch.Tensor, seed: int, particle_offset: int = 0) -> int:
        """Queues one rollout of the pinned host block ``x0_host`` (B, d); returns its ticket.  Does not block unless
        all ``depth`` slots are in flight and un-waited."""
        slot = self._n % self.depth
        if self._busy[slot]:
            raise RuntimeError("HostRolloutStream: wait() for the oldest ticket before submitting another rollout")
        if not x0_host.is_pinned():
            raise N.LrdsError("HostRolloutStream.submit takes pinned host memory (torch.Tensor.pin_memory())")
        comp = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_comp[slot])  # the slot's previous rollout has finished reading its device buffer
            self.x_dev[slot].copy_(x0_host, non_blocking=True)
            self.ev_in[slot].record(self.s_in)
        comp.wait_event(self.ev_in[slot])
        x, rnd = self.simulate(self.x_dev[slot], seed, particle_offset)
        r = rnd.reshape(-1)
        part = torch.empty(8, device=self.device, dtype=torch.float64)
        with torch.cuda.device(self.device):
            N.check(N.lib().lrds_estimator_partials(N.ptr(r), r.numel(), N.ptr(part), N.ptr(self._scratch[slot]),
                                                    N.stream_ptr(self.device)))
        if self.group is not None:
            part = gather_and_merge(part, self.group, on_device=True)  # 64 bytes per rank, merged on the GPU
        self.ev_comp[slot].record(comp)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_comp[slot])
            self.x_host[slot].copy_(x, non_blocking=True)
            self.rnd_host[slot].copy_(rnd, non_blocking=True)
            self.part_host[slot].copy_(part, non_blocking=True)
            self.ev_out[slot].record(self.s_out)
        self._live[slot] = (x, rnd, part)
        self._busy[slot] = True
        self._n += 1
        return self._n - 1

    def wait(self, ticket: int):
        """Blocks until the results of ``ticket`` are in host memory; returns (x_T, rnd, partials) - views of the slot's
        pinned buffers, valid until the slot is submitted again."""
        slot = ticket % self.depth
        self.ev_out[slot].synchronize()
        self._live[slot] = None
        self._busy[slot] = False
        return self.x_host[slot], self.rnd_host[slot], self.part_host[slot]
