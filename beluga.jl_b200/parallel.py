"""Family sharding across GPUs — the reference's only parallelism (DArray of Profiles +
mapreduce, src/profile.jl:27-37,56-75).  One process per GPU; each rank owns a contiguous block of
site patterns; the only exchange is the sum of the per-rank scalars / gradient vectors."""


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of n items owned by `rank` (sizes differ by at most one, the
    larger blocks first — same split as DistributedArrays' default for a vector)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def dist_info():
    """(rank, world) of the default torch.distributed group, (0, 1) if not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def allreduce_sum_(tensor):
    """In-place sum over ranks (NCCL on the tensor's stream for CUDA tensors, gloo for CPU
    tensors); a no-op for a single process."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor
