"""count_confidence / count_delta -- drop-ins for quantpy/tomography/polytopes/utils.py:4-27, evaluated by
qpb_polytope_confidence / qpb_polytope_coverage (csrc/polytope.cu).  Batched variants take a leading trial
axis.  No CPU implementation: a CUDA device is required."""

import ctypes

import numpy as np

from ... import _native as nt


def _flatten_groups(frequencies, n_measurements):
    """The reference broadcasts n_measurements [P] against frequencies [..., P, O]; a leading axis of input states
    (process tomography) simply multiplies the number of POVM groups."""
    f = np.asarray(frequencies, dtype=np.float64)
    n = np.asarray(n_measurements, dtype=np.float64).reshape(-1)
    groups = int(np.prod(f.shape[:-1]))
    reps = groups // n.shape[0]
    return f.reshape(groups, f.shape[-1]), np.tile(n, reps)


def count_confidence(delta, frequencies, n_measurements):
    """Confidence that all true probabilities lie below frequencies + delta (utils.py:4-13)."""
    torch = nt.torch_cuda()
    lib = nt.load_library()
    f, n = _flatten_groups(frequencies, n_measurements)
    fd = nt.to_device(f[None], torch.float64)
    dd = nt.to_device(np.array([delta], dtype=np.float64), torch.float64)
    out = torch.empty((1, 1), dtype=torch.float64, device="cuda")
    nt.check(lib.qpb_polytope_confidence(1, f.shape[0], f.shape[1], nt.ptr(fd), n.ctypes.data_as(ctypes.c_void_p), 1,
                                         nt.ptr(dd), nt.ptr(out), nt.stream_ptr()))
    return float(out.item())


def count_delta(target_cl, frequencies, n_measurements):
    """Smallest shift delta (bisection to 1e-10) whose confidence reaches target_cl (utils.py:16-27)."""
    f, n = _flatten_groups(frequencies, n_measurements)
    return float(count_delta_batch(np.array([target_cl]), f[None], n)[0, 0])


def count_delta_batch(conf_levels, frequencies, n_measurements):
    """count_delta for a batch of trials: frequencies [B, M, O] -> deltas [B, L]."""
    torch = nt.torch_cuda()
    lib = nt.load_library()
    f = np.ascontiguousarray(np.asarray(frequencies, dtype=np.float64))
    n = np.ascontiguousarray(np.asarray(n_measurements, dtype=np.float64).reshape(-1))
    B, M, O = f.shape
    levels = nt.to_device(np.asarray(conf_levels, dtype=np.float64).reshape(-1), torch.float64)
    L = levels.shape[0]
    fd = nt.to_device(f, torch.float64)
    out = torch.empty((B, L), dtype=torch.float64, device="cuda")
    nt.check(lib.qpb_polytope_coverage(B, M, O, None, nt.ptr(fd), n.ctypes.data_as(ctypes.c_void_p), L, nt.ptr(levels),
                                       None, 0, nt.ptr(out), None, nt.stream_ptr()))
    return out.cpu().numpy()
