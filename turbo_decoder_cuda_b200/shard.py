"""Codeblock sharding for multi-GPU runs: codeblocks are independent, so rank g of G
decodes the contiguous index range shard_range(n, G, g) and nothing is exchanged on the data path.
The only collectives are the timing reductions below (max over ranks of a device-timed interval)."""


def shard_range(n, world, rank):
    """[lo, hi) of the codeblocks rank `rank` owns; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _reduce(value, dist, op_name, device):
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op_name))
    return float(t[0])


def max_over_ranks(value, dist, device="cpu"):
    return _reduce(float(value), dist, "MAX", device)


def sum_over_ranks(value, dist, device="cpu"):
    r = _reduce(float(value), dist, "SUM", device)
    return int(r) if isinstance(value, int) else r
