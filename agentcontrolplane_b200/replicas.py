"""Request-level data parallelism for `provider: local`: one engine replica per
GPU, Tasks routed to replicas by a sticky hash of the Task UID (so a Task's KV stays on one GPU
across its turns), NO collective on the data path.  torch.distributed is used only as plumbing
for the benchmark's barrier / max-over-ranks timing (nccl on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import hashlib


def replica_of(task_uid: str, n_replicas: int) -> int:
    """Sticky routing: the same Task always lands on the same replica."""
    h = hashlib.blake2b(task_uid.encode(), digest_size=8).digest()
    return int.from_bytes(h, "little") % max(1, n_replicas)


def shard_tasks(task_uids: list[str], rank: int, world: int) -> list[str]:
    return [u for u in task_uids if replica_of(u, world) == rank]


def aggregate(dist, device, wall_s: float, device_s: float, counts: list[float]):
    """(max over ranks of wall, max over ranks of device time, sum over ranks of counts).
    `dist` is torch.distributed (or None for a single process)."""
    import torch
    t = torch.tensor([wall_s, device_s], dtype=torch.float64, device=device)
    c = torch.tensor(list(counts), dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t[0]), float(t[1]), [float(x) for x in c]
