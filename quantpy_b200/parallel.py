"""Multi-GPU plumbing for the bootstrap: shard the resample index range over ranks and collect the
per-sample statistics with ONE all-gather.  There is no other exchange: resamples are
independent, and the Philox counter is the GLOBAL sample index, so the gathered vector is identical
for any world size."""

import numpy as np


def world():
    """(rank, world_size) of the default process group, (0, 1) when torch.distributed is not initialised."""
    try:
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_items, rank, world_size):
    """Contiguous slice [lo, hi) of range(n_items) owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_seed(seed):
    """All ranks must sample with rank 0's key."""
    rank, size = world()
    if size == 1:
        return seed
    import torch
    import torch.distributed as dist

    device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([seed if rank == 0 else 0], dtype=torch.int64, device=device)
    dist.broadcast(t, src=0)
    return int(t.item())


def all_gather_concat(local, n_total):
    """Concatenate every rank's 1-D tensor in rank order (slices from shard_bounds) -> length n_total.

    Uses a single all_gather on equal-sized (padded) buffers: N*8 bytes in total over NVLink."""
    rank, size = world()
    if size == 1:
        return local
    import torch
    import torch.distributed as dist

    width = -(-int(n_total) // size)
    padded = torch.zeros((width,), dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    gathered = torch.empty((size * width,), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded)
    pieces = []
    for r in range(size):
        lo, hi = shard_bounds(n_total, r, size)
        pieces.append(gathered[r * width: r * width + (hi - lo)])
    return torch.cat(pieces)


def quantile_function(dist_values, presorted=False):
    """Sorted distances -> interp1d(linspace(0, 1, N), dist) as quantpy/tomography/interval.py:610-612."""
    from scipy.interpolate import interp1d

    ordered = np.asarray(dist_values, dtype=np.float64)
    if not presorted:
        ordered = np.sort(ordered)
    return interp1d(np.linspace(0, 1, len(ordered)), ordered, assume_sorted=True)
