"""Host-side sharding of independent transforms across ranks (one process per GPU).

Polynomials -- and RNS limbs, each with its own modulus and tables -- are independent units
(SURVEY.md section 8e), so a job is partitioned contiguously across ranks with NO data-path collective;
torch.distributed is only used for the barrier and the max-over-ranks timing reduction.
"""


def shard_range(total, rank, world):
    """Contiguous [begin, end) of `total` units owned by `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def limb_owner(limb, limbs, world):
    """Rank that owns RNS limb `limb` under shard_range."""
    for r in range(world):
        b, e = shard_range(limbs, r, world)
        if b <= limb < e:
            return r
    raise ValueError("limb out of range")


def reduce_max(value, dist=None):
    """Max over ranks of a python float (the whole-job time of a step is the slowest rank's)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value, dist=None):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
