#!/usr/bin/env python
"""Sharded LV training check, run under torchrun with N ranks (one per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_2gpu_check.py
Every rank integrates its shard of the batch (Philox noise keyed by the global particle index), the variance weights
come from one all_reduce of three sums and the parameter gradients from one all_reduce; rank 0 then recomputes loss
and gradient of the WHOLE batch on its own GPU and compares."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tests import cases as T  # noqa: E402
from tests.product_builders import Built  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    case = T.case_ei_many_modes(K=50, B=1024)
    x0 = T.initial_state(case)
    built = Built(case, dev, "f16x3")
    params = dict(built.ctrl.named_parameters())
    n = x0.shape[0] // world
    loss, _ = built.train_loss(x0[rank * n:(rank + 1) * n], None, seed=99, particle_offset=rank * n, group=dist.group.WORLD)
    loss.backward()
    sharded = {k: p.grad.detach().clone() for k, p in params.items()}
    ok = True
    if rank == 0:
        for p in params.values():
            p.grad = None
        whole, _ = built.train_loss(x0, None, seed=99, particle_offset=0)
        whole.backward()
        worst = max(((sharded[k] - p.grad).abs().max() / p.grad.abs().max().clamp(min=1e-12)).item() for k, p in params.items())
        rel = abs(loss.item() - whole.item()) / abs(whole.item())
        ok = worst < 1e-4 and rel < 1e-5
        print(f"world {world}: sharded loss {loss.item():.6f} vs whole batch {whole.item():.6f} (rel {rel:.1e}); "
              f"worst gradient tensor error {worst:.1e} -> {'OK' if ok else 'MISMATCH'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
