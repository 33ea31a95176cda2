"""ctypes binding of libquantpy_b200.so (include/quantpy_b200.h).

PyTorch is used only for plumbing: device memory (tensors), the current CUDA stream and
torch.distributed.  There is NO CPU fallback: every entry point raises if the shared library or
a CUDA device is missing.
"""

import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libquantpy_b200.so")

DIST_KINDS = {"hs": 0, "trace": 1, "if": 2}
METHODS = {"lin": 0, "mle": 1}
INITS = {"lin": 0, "mixed": 1}

_lib = None
_lock = threading.Lock()

c_i32p = ctypes.c_void_p
_vp = ctypes.c_void_p
_int = ctypes.c_int
_dbl = ctypes.c_double
_u64 = ctypes.c_uint64

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "qpb_abi_version": (_int, []),
    "qpb_last_error": (ctypes.c_char_p, []),
    "qpb_launch_count": (ctypes.c_int64, []),
    "qpb_debug_set_trace": (_int, [_vp, ctypes.c_size_t]),
    "qpb_lin_project_ordered": (_int, [_vp, _int, _vp, _vp, _vp, _vp]),
    "qpb_mle_rrr_ordered": (_int, [_vp, _int, _vp, _vp, _vp, _int, _dbl, _vp, _vp, _vp]),
    "qpb_identity_order": (_int, [_int, _vp, _vp]),
    "qpb_reset_launch_count": (None, []),
    "qpb_set_option": (_int, [_int, _int]),
    "qpb_get_option": (_int, [_int]),
    "qpb_fp64_fma_probe": (_int, [ctypes.c_int64, _vp, ctypes.POINTER(_dbl), _vp]),
    "qpb_smem_probe": (_int, [ctypes.c_int64, _vp, ctypes.POINTER(_dbl), _vp]),
    "qpb_state_plan_create": (_int, [ctypes.POINTER(_vp), _int, _int, _vp, _vp, _vp]),
    "qpb_state_plan_destroy": (_int, [_vp]),
    "qpb_povm_probs": (_int, [_int, _int, _int, _vp, _vp, _dbl, _int, _vp, _vp]),
    "qpb_multinomial": (_int, [_int, _int, _int, _vp, _int, _vp, _u64, _u64, _vp, _vp]),
    "qpb_lin_project": (_int, [_vp, _int, _vp, _int, _vp, _vp]),
    "qpb_mle_rrr": (_int, [_vp, _int, _vp, _vp, _int, _dbl, _vp, _vp, _vp]),
    "qpb_mle_variant": (_int, [_vp]),
    "qpb_distance": (_int, [_int, _int, _vp, _vp, _int, _vp, _vp]),
    "qpb_bootstrap_state_workspace": (ctypes.c_size_t, [_vp, _int, _int, _int]),
    "qpb_bootstrap_state": (_int, [_vp, _int, _int, _int, _vp, _vp, _u64, _u64, _int, _int, _int, _int, _dbl,
                                   _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qpb_choi_from_states": (_int, [_int, _int, _int, _vp, _vp, _vp, _vp]),
    "qpb_cptp_project_if_needed": (_int, [_int, _int, _vp, _int, _dbl, _dbl, _vp, _vp, _vp]),
    "qpb_polytope_coverage": (_int, [_int, _int, _int, _vp, _vp, _vp, _int, _vp, _vp, _int, _vp, _vp, _vp]),
    "qpb_polytope_confidence": (_int, [_int, _int, _int, _vp, _vp, _int, _vp, _vp, _vp]),
    "qpb_l2_moments": (_int, [_int, _int, _int, _vp, _vp, _dbl, _vp, _vp, _vp]),
    "qpb_sort_f64": (_int, [ctypes.c_longlong, _vp, _vp, _vp]),
    "qpb_merge_sorted_runs": (_int, [_int, _vp, _vp, _vp, _vp, _vp]),
    "qpb_quantiles_host": (_int, [ctypes.c_longlong, _vp, _int, _vp, _vp, _vp]),
    "qpb_bootstrap_state_interval": (_int, [_vp, _int, _int, _int, _vp, _vp, _vp, _vp, _u64, _u64, _int, _int, _int,
                                            _int, _dbl, _int, _int, _vp, _vp, _vp, _vp, _vp]),
    "qpb_mhmc_state": (_int, [_vp, _int, _int, _int, _int, _dbl, _vp, _int, _vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp,
                              _vp]),
    "qpb_process_plan_create": (_int, [ctypes.POINTER(_vp), _int, _int, _int, _vp, _vp]),
    "qpb_process_plan_destroy": (_int, [_vp]),
    "qpb_lifp_cptp": (_int, [_vp, _int, _vp, _int, _int, _dbl, _vp, _vp, _vp]),
    "qpb_cptp_project": (_int, [_int, _int, _vp, _int, _dbl, _vp, _vp, _vp]),
}


# QPB_OPT_* of include/quantpy_b200.h (kernel-selection toggles for tests and profiling)
OPTIONS = {"NO_TAIL_MERGE": 0, "NO_HS_FUSION": 1, "NO_PAULI_KERNEL": 2, "NO_CONST_KERNEL": 3, "NO_AXIS_KERNEL": 4,
           "NO_DMMA_GEMM": 5, "NO_ROW_JACOBI": 6, "NO_PACKED_JACOBI": 7, "NO_LIN_SMALL": 8, "SAMPLER": 9,
           "MLE_BLOCKS_PER_SM": 10, "MLE_LANES": 11, "NO_TILED_MLE": 12, "MLE_PARK_AGE": 13, "MLE_PARK_LIVE": 14,
           "MLE_W_WARPS": 15, "MLE_PARK_PLATEAU": 16, "NO_TMA_GEMM": 17, "MLE_TAIL_POLL": 18, "MLE_TAIL_AGE": 19,
           "MLE_ADOPT": 20, "MLE_MERGE": 21, "NO_MLE_ORDER": 22,
           "MLE_PARK_AGE_LO": 23, "MLE_PARK_AGE_PCT": 24, "MLE_PARK_AGE_END": 25, "MLE_PARK_AGE_PCT2": 26,
           "MLE_REFILL_MIN": 27, "NO_WARM_JACOBI": 28, "MLE_SINGLE_WARPS": 29,
           "SAMPLER_NO_PREFILTER": 30, "SAMPLER_EXACT_EVERY": 31, "SAMPLER_THREADS": 32, "NO_SAMPLE_SORT": 33,
           "SAMPLER_LANES": 34}
SAMPLERS = {"auto": 0, "alias": 1, "binomial": 2}


class NativeError(RuntimeError):
    pass


def set_option(name, value):
    """Set a kernel-selection toggle (tests / profiling); returns the previous value."""
    lib = load_library()
    old = lib.qpb_get_option(OPTIONS[name])
    check(lib.qpb_set_option(OPTIONS[name], int(value)))
    return old


class option:
    """Context manager: `with option("NO_TAIL_MERGE", 1): ...` restores the previous value on exit."""

    def __init__(self, name, value):
        self.name, self.value = name, value

    def __enter__(self):
        self.old = set_option(self.name, self.value)
        return self

    def __exit__(self, *exc):
        set_option(self.name, self.old)
        return False


def load_library():
    """Load the shared library (no GPU needed).  Raises NativeError if it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise NativeError(
                    f"{LIB_PATH} is missing: build it with `make -C quantpy_b200/csrc` "
                    "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
                    "quantpy_b200 has no CPU fallback."
                )
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            if lib.qpb_abi_version() != 1:
                raise NativeError("libquantpy_b200.so ABI version mismatch")
            _lib = lib
    return _lib


_torch = None


def torch_cuda():
    """Return the torch module after checking that a CUDA device is usable (checked once: the probe costs a few
    microseconds and every entry point of the package passes through here)."""
    global _torch
    if _torch is None:
        import torch

        if not torch.cuda.is_available():
            raise NativeError("quantpy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
        _torch = torch
    return _torch


def check(rc):
    if rc != 0:
        msg = load_library().qpb_last_error()
        raise NativeError(f"libquantpy_b200 error {rc}: {msg.decode() if msg else '?'}")


def stream_ptr():
    torch = torch_cuda()
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def to_device(array, dtype):
    """Host array -> contiguous device tensor of the given torch dtype."""
    torch = torch_cuda()
    a = np.ascontiguousarray(array)
    return torch.from_numpy(a).to(device="cuda", dtype=dtype, non_blocking=False).contiguous()


def complex_to_host(t):
    """Device float64 tensor [..., 2] of (re, im) pairs -> complex128 numpy array."""
    return np.ascontiguousarray(t.cpu().numpy()).view(np.complex128)[..., 0]


def complex_to_device(array):
    """complex numpy array -> device float64 tensor with a trailing (re, im) axis."""
    torch = torch_cuda()
    a = np.ascontiguousarray(np.asarray(array, dtype=np.complex128))
    return torch.from_numpy(a.view(np.float64).reshape(a.shape + (2,))).cuda()
