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

    gathered, width = _all_gather_padded(local, n_total)
    if int(n_total) == size * width:  # equal shards: the gathered buffer already is the concatenation
        return gathered
    pieces = []
    for r in range(size):
        lo, hi = shard_bounds(n_total, r, size)
        pieces.append(gathered[r * width: r * width + (hi - lo)])
    return torch.cat(pieces)


def _all_gather_padded(local, n_total):
    """ONE all_gather_into_tensor of every rank's shard, padded to the common width -> (buffer [size*width], width)."""
    import torch
    import torch.distributed as dist

    rank, size = world()
    width = -(-int(n_total) // size)
    gathered = torch.empty((size * width,), dtype=local.dtype, device=local.device)
    if local.numel() == width:
        dist.all_gather_into_tensor(gathered, local.contiguous())
    else:
        padded = torch.zeros((width,), dtype=local.dtype, device=local.device)
        padded[: local.numel()] = local
        dist.all_gather_into_tensor(gathered, padded)
    return gathered, width


def gather_sorted(local, n_total, presorted=False):
    """The sorted vector of all ranks' values (what the reference gets from `dist.sort()`, interval.py:610): every
    rank sorts its own shard, ONE all-gather collects the sorted shards, and a counting merge of the world_size
    runs (qpb_merge_sorted_runs, one launch) replaces a second full sort on every rank.  CUDA tensors only; the
    CPU (gloo) path used by the host-logic tests sorts the concatenation."""
    rank, size = world()
    if not local.is_cuda:
        import torch

        return torch.sort(all_gather_concat(local, n_total)).values  # (a presorted shard changes nothing here)
    import ctypes

    import torch

    from . import _native as nt
    from . import engine

    mine = local if presorted else engine.sort_f64(local)
    if size == 1:
        return mine
    gathered, width = _all_gather_padded(mine, n_total)
    lens = np.zeros(size, dtype=np.int32)
    starts = np.zeros(size, dtype=np.int64)
    for r in range(size):
        lo, hi = shard_bounds(n_total, r, size)
        lens[r], starts[r] = hi - lo, r * width
    out = torch.empty((int(n_total),), dtype=local.dtype, device=local.device)
    nt.check(nt.load_library().qpb_merge_sorted_runs(size, lens.ctypes.data_as(ctypes.c_void_p),
                                                     starts.ctypes.data_as(ctypes.c_void_p), nt.ptr(gathered),
                                                     nt.ptr(out), nt.stream_ptr()))
    return out


# bytes that crossed PCIe through this module since import (bench.py reads the difference around a call)
TRAFFIC = {"d2h": 0, "h2d": 0}

_PINNED = {}


def _pinned(dtype, numel):
    """Cached page-locked staging buffer (grow-only per dtype)."""
    import torch

    buf = _PINNED.get(dtype)
    if buf is None or buf.numel() < numel:
        buf = _PINNED[dtype] = torch.empty((max(int(numel), 4096),), dtype=dtype, pin_memory=True)
    return buf[:numel]


def to_host_pinned(tensor):
    """Device tensor -> NumPy array through a cached page-locked staging buffer (pageable copies of a few MB run
    at a fraction of the PCIe rate and dominated the multi-GPU end-to-end time)."""
    import torch

    if not tensor.is_cuda:
        return tensor.numpy().copy()
    buf = _pinned(tensor.dtype, tensor.numel()).view(tensor.shape)
    buf.copy_(tensor, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    TRAFFIC["d2h"] += tensor.numel() * tensor.element_size()
    return buf.numpy().copy()


class QuantileFunction:
    """Linear interpolant of the sorted bootstrap distances on the grid linspace(0, 1, N) -- the object the
    reference builds with scipy.interpolate.interp1d(conf_levels, dist) (quantpy/tomography/interval.py:610-612).
    Same call semantics (vectorised, ValueError outside [0, 1]) and the same `.x` / `.y` attributes.

    The sorted values may stay ON THE DEVICE: a call then sends the requested levels and receives their quantiles
    (a few KB, qpb_quantiles_host), and nothing of size N crosses PCIe unless `.y` (or the interval's `.dist`) is
    asked for."""

    def __init__(self, sorted_values):
        self._dev = None
        self._y = None
        if hasattr(sorted_values, "is_cuda") and sorted_values.is_cuda:
            self._dev = sorted_values
            self._n = int(sorted_values.numel())
        else:
            if hasattr(sorted_values, "numpy"):
                sorted_values = sorted_values.numpy()
            self._y = np.asarray(sorted_values, dtype=np.float64)
            self._n = len(self._y)
        self._x = None
        self._primed = None

    @property
    def y(self):
        if self._y is None:
            self._y = to_host_pinned(self._dev)
        return self._y

    @property
    def x(self):
        if self._x is None:
            self._x = np.linspace(0, 1, self._n)
        return self._x

    def __call__(self, levels):
        levels = np.asarray(levels, dtype=np.float64)
        if levels.size:  # (min / max: the same verdicts as any(levels < 0) / any(levels > 1), without temporaries)
            if levels.min() < 0:
                raise ValueError("A value in x_new is below the interpolation range.")
            if levels.max() > 1:
                raise ValueError("A value in x_new is above the interpolation range.")
        n = self._n
        if self._y is None:
            # device-resident: the levels go up, the quantiles come back (qpb_quantiles_host evaluates the expression
            # below in the same IEEE operations); a call with the levels the interval was set up with (the same bytes)
            # is answered from what that call already brought back
            if self._primed is not None and self._primed[0] == levels.shape and self._primed[1] == levels.tobytes():
                return self._primed[2].copy()
            from . import engine

            TRAFFIC["h2d"] += levels.size * 8
            TRAFFIC["d2h"] += levels.size * 8
            return engine.quantiles_host(self._dev, levels)
        if n == 1:
            return np.full(levels.shape, self._y[0])
        pos = levels * (n - 1)
        lo = np.minimum(np.floor(pos).astype(np.int64), n - 2)
        frac = pos - lo
        y_lo, y_hi = self._y[lo], self._y[lo + 1]
        return y_lo + (y_hi - y_lo) * frac

    def prime(self, levels, values):
        """Remember the quantiles a fused set-up call already returned for `levels`."""
        levels = np.asarray(levels, dtype=np.float64)
        self._primed = (levels.shape, levels.tobytes(), np.array(values, dtype=np.float64).reshape(levels.shape))


def quantile_function(dist_values, presorted=False):
    """Sorted distances -> interpolant of (linspace(0, 1, N), dist) as quantpy/tomography/interval.py:610-612.
    A CUDA tensor stays on the device (sorted there if needed)."""
    if hasattr(dist_values, "is_cuda") and dist_values.is_cuda:
        if not presorted:
            from . import engine

            dist_values = engine.sort_f64(dist_values)
        return QuantileFunction(dist_values)
    ordered = np.asarray(dist_values, dtype=np.float64)
    if not presorted:
        ordered = np.sort(ordered)
    return QuantileFunction(ordered)
