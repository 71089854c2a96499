"""
Batch sharding across the GPUs of one box.  Every problem of the hot path is independent, so the batch axis
is split contiguously by rank and there is NO collective on the data path; the only communication is the optional final
gather of results (costs, trajectories), done here with `torch.distributed.all_gather` (NCCL over NVLink on the GPU box,
gloo in the CPU tests).  Host logic only -- no arithmetic.
"""
import torch
import torch.distributed as dist


def shard_range(Bsz, rank, world):
    """Contiguous split of range(Bsz): rank r owns [lo, hi); sizes differ by at most one, earlier ranks get the extra."""
    base, rem = divmod(int(Bsz), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t, rank, world, dim=0):
    """This rank's slice of a batched array (a view, no copy)."""
    lo, hi = shard_range(t.shape[dim], rank, world)
    return t.narrow(dim, lo, hi - lo) if isinstance(t, torch.Tensor) else t[(slice(None),) * dim + (slice(lo, hi),)]


def gather(local, Bsz, group=None):
    """Concatenate per-rank results (ragged shards allowed) in rank order on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    sizes = [shard_range(Bsz, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = local.new_zeros((maxn,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)
