"""Shot sharding over the GPUs of one box.

Shots are independent, so rank r of R simply takes the global shot ids [r*N/R, (r+1)*N/R); the Philox
counter is the global shot id, which makes every counter independent of R.  The only exchange is one
all-reduce(sum) of the integer counters at the end (NCCL over NVLink on GPUs, gloo in CPU tests).
"""
import os


def world():
    """(rank, world_size) from torch.distributed when initialised, else from the torchrun env, else (0, 1)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(n, rank, world_size):
    """Contiguous, balanced split of range(n): returns (first, count) for `rank`."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n, world_size)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def allreduce_counts(values, device=None):
    """Sum a list of Python ints over all ranks (no-op when torch.distributed is not initialised)."""
    try:
        import torch
        import torch.distributed as dist
    except ImportError:
        return list(values)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(values)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(v) for v in t.tolist()]
