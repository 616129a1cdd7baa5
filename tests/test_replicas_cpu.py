"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the sticky Task routing and the
max-over-ranks / sum-over-ranks aggregation bench.py uses (no GPU, no collective on the data path)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from agentcontrolplane_b200 import host
from agentcontrolplane_b200.replicas import aggregate, replica_of, shard_tasks


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uids = [f"uid-{i}" for i in range(101)]
    mine = shard_tasks(uids, rank, world)
    # each rank reconciles ITS Tasks through the CPU path (stub completion server, config 0)
    with host.StubServer() as srv:
        r = host.hostsim_run({"tasks": len(mine), "workers": 2, "provider": "openai", "model": "m", "baseURL": srv.base_url})
    wall, dev, counts = aggregate(dist, "cpu", 1.0 + rank, 0.5 * (rank + 1), [float(r["reconciles"]), float(len(mine))])
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        q.put((wall, dev, counts, gathered, r["final_phases"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_replicas_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    wall, dev, counts, gathered, phases = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert wall == 2.0 and dev == 1.0                       # max over ranks
    assert counts == [101.0, 101.0]                         # every Task reconciled exactly once
    flat = sorted(gathered[0] + gathered[1])
    assert flat == sorted(f"uid-{i}" for i in range(101))   # a partition: no Task on two replicas
    assert set(gathered[0]).isdisjoint(gathered[1]) and min(len(gathered[0]), len(gathered[1])) > 30
    assert list(phases) == ["FinalAnswer"]


def test_routing_is_sticky_and_balanced():
    assert all(replica_of(f"t{i}", 8) == replica_of(f"t{i}", 8) for i in range(50))
    counts = [0] * 8
    for i in range(8000):
        counts[replica_of(f"task-{i}", 8)] += 1
    assert min(counts) > 800 and max(counts) < 1200
    assert replica_of("x", 1) == 0
    w, d, c = aggregate(None, "cpu", 1.5, 0.25, [3.0])
    assert (w, d, c) == (1.5, 0.25, [3.0])
