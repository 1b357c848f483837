"""Shot sharding over the GPUs of one box.

Shots are independent, so rank r of R simply takes the global shot ids [r*N/R, (r+1)*N/R); the Philox
counter is the global shot id, which makes every counter independent of R.  The only exchange is one
all-reduce(sum) of the integer counters at the end (NCCL over NVLink on GPUs, gloo in CPU tests).
"""
import os


_warned = False


def world():
    """(rank, world_size) of the initialised torch.distributed process group, else (0, 1).

    Sharding and the counter all-reduce use the SAME condition: a process that was launched by torchrun but never
    called init_process_group simulates the whole shot range itself (every rank returns the full, correct counters;
    a warning says so) instead of returning a silent fraction of them."""
    global _warned
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not _warned:
        _warned = True
        import sys
        print("surface_17_iqt_b200: WORLD_SIZE > 1 but torch.distributed is not initialised -- this process runs every "
              "shot itself (call torch.distributed.init_process_group to shard the shots over the ranks)", file=sys.stderr)
    return 0, 1


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


def _group_device(dist):
    import torch
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def broadcast_seed(seed):
    """Rank 0's 64-bit seed on every rank (each rank draws its own from Python's `random`, whose state need not agree
    between ranks; the counters are independent of the sharding only if all ranks use one Philox key)."""
    try:
        import torch
        import torch.distributed as dist
    except ImportError:
        return seed
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return seed
    t = torch.tensor([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=torch.int64, device=_group_device(dist))
    dist.broadcast(t, src=0)
    lo, hi = [int(v) for v in t.tolist()]
    return lo | (hi << 32)


def gather_lists(values):
    """Concatenate one list of ints per rank in rank order (run-til-fail: rounds_til_fail_list of every rank's runs).
    Accepts and returns a numpy array when given one (large run counts never become Python ints on the way)."""
    import numpy as np
    as_array = isinstance(values, np.ndarray)
    try:
        import torch
        import torch.distributed as dist
    except ImportError:
        return values if as_array else list(values)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return values if as_array else list(values)
    dev = _group_device(dist)
    n = torch.tensor([len(values)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(dist.get_world_size())]
    dist.all_gather(sizes, n)
    sizes = [int(v.item()) for v in sizes]
    pad = max(sizes) if sizes else 0
    mine = torch.zeros(pad, dtype=torch.int64, device=dev)
    if len(values):
        mine[:len(values)] = torch.as_tensor(np.asarray(values, dtype=np.int64), device=dev)
    parts = [torch.zeros(pad, dtype=torch.int64, device=dev) for _ in sizes]
    dist.all_gather(parts, mine)
    out = np.concatenate([part[:sz].cpu().numpy() for sz, part in zip(sizes, parts)]) if sizes else np.zeros(0, np.int64)
    return out if as_array else [int(v) for v in out]


def gather_to_root(values):
    """Rank 0 receives the concatenation (rank order) of one int64 array per rank; the other ranks get an empty array."""
    import numpy as np
    values = np.asarray(values, dtype=np.int64)
    try:
        import torch
        import torch.distributed as dist
    except ImportError:
        return values
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return values
    dev = _group_device(dist)
    rank, world = dist.get_rank(), dist.get_world_size()
    n = torch.tensor([len(values)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(v.item()) for v in sizes]
    pad = max(sizes) if sizes else 0
    mine = torch.zeros(pad, dtype=torch.int64, device=dev)
    if len(values):
        mine[:len(values)] = torch.as_tensor(values, device=dev)
    parts = [torch.zeros(pad, dtype=torch.int64, device=dev) for _ in sizes] if rank == 0 else None
    dist.gather(mine, parts, dst=0)
    if rank != 0:
        return np.zeros(0, np.int64)
    return np.concatenate([part[:sz].cpu().numpy() for sz, part in zip(sizes, parts)])


def broadcast_float(value):
    """Rank 0's float (None allowed on the others) on every rank."""
    try:
        import torch
        import torch.distributed as dist
    except ImportError:
        return value
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([float(value) if value is not None else 0.0], dtype=torch.float64, device=_group_device(dist))
    dist.broadcast(t, src=0)
    return float(t.item())
