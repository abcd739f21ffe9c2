"""Multi-GPU plumbing: QPs are independent, so a batch shards by contiguous index range over the
GPUs of one box with no data-path collective. torch.distributed is used only for
the barrier and the max-over-ranks timing reduction of the benchmark harness."""


def shard_range(batch, rank, world_size):
    """Contiguous range [lo, hi) of the global batch owned by `rank`: sizes differ by at most one."""
    base, rem = divmod(int(batch), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def all_shards(batch, world_size):
    return [shard_range(batch, r, world_size) for r in range(world_size)]


def reduce_max_time(seconds, device=None):
    """MAX over ranks of a per-rank duration (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(seconds)
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
