"""Multi-GPU sharding of the hot path.

Clusters -- and the <=1000-read batches inside them -- are independent units (the reference runs one
`python main` subprocess per batch, isONform_parallel:250-311, and re-joins batches per cluster,
batch_merging_parallel.py:227-276).  So: one process per GPU, whole clusters assigned to ranks, no
collective on the data path.  torch.distributed is used only to agree on counters/timings.
"""
import os


def batch_sizes(n_reads, max_seqs=1000):
    """Batch sizes isONform_parallel produces for a cluster of n_reads (isONform_parallel:96-209:
    chunks of max_seqs reads, the remainder in a last chunk)."""
    full, rest = divmod(int(n_reads), int(max_seqs))
    return [max_seqs] * full + ([rest] if rest else [])


def cluster_cost(n_reads, max_seqs=1000):
    """Work estimate: the all-pairs stages are quadratic in the batch size, and the cross-batch merge is
    quadratic in the number of batches."""
    bs = batch_sizes(n_reads, max_seqs)
    return sum(b * b for b in bs) + (len(bs) * (len(bs) - 1) // 2) * max_seqs


def assign_clusters(cluster_reads, world_size, max_seqs=1000):
    """Longest-processing-time greedy: {cluster_id: n_reads} -> list (per rank) of cluster ids.
    Deterministic: ties broken by cluster id, so every rank computes the same plan without talking."""
    order = sorted(cluster_reads, key=lambda c: (-cluster_cost(cluster_reads[c], max_seqs), str(c)))
    load = [0] * world_size
    plan = [[] for _ in range(world_size)]
    for c in order:
        r = min(range(world_size), key=lambda x: (load[x], x))
        plan[r].append(c)
        load[r] += cluster_cost(cluster_reads[c], max_seqs)
    return plan


def my_clusters(cluster_reads, rank=None, world_size=None, max_seqs=1000):
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world_size = int(os.environ.get("WORLD_SIZE", "1")) if world_size is None else world_size
    return assign_clusters(cluster_reads, world_size, max_seqs)[rank]


def split_pairs(n_pairs, rank, world_size):
    """Even split of one giant pair list (a single huge cluster / the pair microbenchmark): pairs are
    independent.  Returns (start, stop)."""
    base, extra = divmod(int(n_pairs), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def reduce_counters(counters, op="sum"):
    """All-reduce a {name: number} dict across ranks (nccl or gloo); identity when not initialised."""
    try:
        import torch
        import torch.distributed as dist
    except ImportError:                                  # pragma: no cover
        return dict(counters)
    if not (dist.is_available() and dist.is_initialized()):
        return dict(counters)
    names = sorted(counters)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(counters[n]) for n in names], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return {n: float(v) for n, v in zip(names, t.tolist())}
