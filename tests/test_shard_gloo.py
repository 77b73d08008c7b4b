"""CPU: the N>1 path -- cluster sharding and counter reduction -- with world_size 2 over gloo."""
import os
import socket

import pytest

from isonform_b200 import shard


def test_assign_clusters_is_a_partition_and_balanced():
    sizes = {c: 37 * ((c * 7919) % 83 + 1) for c in range(200)}
    for world in (1, 2, 4, 8):
        plan = shard.assign_clusters(sizes, world)
        flat = [c for p in plan for c in p]
        assert sorted(flat) == sorted(sizes) and len(plan) == world
        load = [sum(shard.cluster_cost(sizes[c]) for c in p) for p in plan]
        assert max(load) <= 1.05 * (sum(load) / world) + max(shard.cluster_cost(v) for v in sizes.values())
        assert plan == shard.assign_clusters(dict(reversed(list(sizes.items()))), world)     # order independent


def test_batch_sizes_and_pair_split():
    assert shard.batch_sizes(1000) == [1000] and shard.batch_sizes(1001) == [1000, 1] and shard.batch_sizes(5) == [5]
    for n, w in ((10, 3), (7, 8), (499500, 8), (0, 2)):
        spans_ = [shard.split_pairs(n, r, w) for r in range(w)]
        assert spans_[0][0] == 0 and spans_[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans_, spans_[1:]))
        assert max(b - a for a, b in spans_) - min(b - a for a, b in spans_) <= 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes = {c: 100 + 50 * c for c in range(11)}
    mine = shard.my_clusters(sizes)                                  # reads RANK / WORLD_SIZE
    work = {"pairs": sum(sizes[c] * (sizes[c] - 1) // 2 for c in mine), "clusters": len(mine), "ms": 10.0 * (rank + 1)}
    tot = shard.reduce_counters({"pairs": work["pairs"], "clusters": work["clusters"]}, "sum")
    mx = shard.reduce_counters({"ms": work["ms"]}, "max")
    dist.barrier()
    q.put((rank, mine, tot, mx))
    dist.destroy_process_group()


def test_two_ranks_gloo_shard_and_reduce():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sizes = {c: 100 + 50 * c for c in range(11)}
    assert sorted(res[0][1] + res[1][1]) == list(range(11)) and not set(res[0][1]) & set(res[1][1])
    want_pairs = sum(v * (v - 1) // 2 for v in sizes.values())
    for _rank, _mine, tot, mx in res:
        assert tot == {"clusters": 11.0, "pairs": float(want_pairs)} and mx == {"ms": 20.0}


def test_reduce_counters_identity_without_process_group():
    assert shard.reduce_counters({"a": 3, "b": 4.5}) == {"a": 3, "b": 4.5}
