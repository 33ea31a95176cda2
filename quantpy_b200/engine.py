"""Device-side engine: thin Python wrappers that hand device pointers to libquantpy_b200.so.

Everything here is plumbing -- tensors as device buffers, the current CUDA stream, plan caching.
All arithmetic on the hot path happens inside the CUDA kernels (csrc/).  No function in this module
has a CPU implementation; each raises NativeError when the library or a GPU is missing.
"""

import ctypes
import hashlib

import numpy as np

from . import _native as nt
from .routines import _left_inv


def next_seed():
    """Draw a 63-bit Philox key from NumPy's global legacy RNG, so that `np.random.seed(s)` makes
    simulated experiments reproducible exactly where it does for the reference (state.py:112 uses the
    same global stream; the stream itself is not reproduced)."""
    return int(np.random.randint(0, 2**63 - 1, dtype=np.int64))


def weighted_povm(povm_matrix, n_measurements):
    """(K, D) shot-weighted POVM rows (quantpy/tomography/state.py:193-196)."""
    povm_matrix = np.asarray(povm_matrix, dtype=np.float64)
    n = np.asarray(n_measurements, dtype=np.float64).reshape(-1)
    return np.reshape(povm_matrix * n[:, None, None] / n.sum(), (-1, povm_matrix.shape[-1]))


def sample_counts(probs, n_samples, P, O, n_shots, seed, offset=0):
    """qpb_multinomial: counts [B, P, O] int32 on the device.  probs: device [P*O] (shared by all
    samples) or [B, P*O]; n_shots: host int array [P]."""
    torch = nt.torch_cuda()
    lib = nt.load_library()
    batched = int(probs.dim() == 2 and probs.shape[0] == n_samples and n_samples > 1)
    counts = torch.empty((n_samples, P, O), dtype=torch.int32, device="cuda")
    shots = np.ascontiguousarray(np.asarray(n_shots, dtype=np.int32).reshape(-1))
    if len(shots) != P:
        raise ValueError("Wrong length for argument `n_measurements`")
    nt.check(lib.qpb_multinomial(n_samples, P, O, nt.ptr(probs.contiguous()), batched,
                                 shots.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(seed),
                                 ctypes.c_uint64(offset), nt.ptr(counts), nt.stream_ptr()))
    return counts


class StatePlan:
    """Device tables for one (POVM tensor, shot vector): wraps qpb_state_plan."""

    def __init__(self, povm_matrix, n_measurements, need_lin=True):
        torch = nt.torch_cuda()
        lib = nt.load_library()
        povm_matrix = np.asarray(povm_matrix, dtype=np.float64)
        self.P, self.O, self.D = povm_matrix.shape
        self.K = self.P * self.O
        self.n_qubits = int(round(np.log2(self.D) / 2))
        self.d = 2**self.n_qubits
        if 4**self.n_qubits != self.D or not 1 <= self.n_qubits <= 4:
            raise ValueError("POVM rows must have length 4^n with 1 <= n <= 4")
        A = weighted_povm(povm_matrix, n_measurements)
        self.A_dev = nt.to_device(A, torch.float64)
        self.L_dev = nt.to_device(_left_inv(A), torch.float64) if need_lin else None
        self.M_dev = nt.to_device(povm_matrix.reshape(self.K, self.D), torch.float64)
        self.n_shots = np.asarray(np.rint(np.asarray(n_measurements, dtype=np.float64)), dtype=np.int32).reshape(-1)
        handle = ctypes.c_void_p()
        nt.check(lib.qpb_state_plan_create(ctypes.byref(handle), self.n_qubits, self.K, nt.ptr(self.A_dev),
                                           nt.ptr(self.L_dev), nt.stream_ptr()))
        self.handle = handle
        self._lib = lib

    def __del__(self):
        handle = getattr(self, "handle", None)
        if handle:
            try:
                self._lib.qpb_state_plan_destroy(handle)
            except Exception:
                pass
            self.handle = None

    # -- kernels ---------------------------------------------------------------------------------
    def probabilities(self, bloch):
        """clip(2^n M.r, 0, 1) for a batch of Bloch vectors -> device tensor [B, K] (k1)."""
        torch = nt.torch_cuda()
        r = nt.to_device(np.atleast_2d(np.asarray(bloch, dtype=np.float64)), torch.float64)
        B = r.shape[0]
        p = torch.empty((B, self.K), dtype=torch.float64, device="cuda")
        nt.check(self._lib.qpb_povm_probs(self.K, self.D, B, nt.ptr(self.M_dev), nt.ptr(r), float(self.d), 1,
                                          nt.ptr(p), nt.stream_ptr()))
        return p

    def sample(self, probs, n_samples, seed, offset=0):
        """Multinomial counts [B, P, O] int32 on the device (k2).  probs: device [K] or [B, K]."""
        return sample_counts(probs, n_samples, self.P, self.O, self.n_shots, seed, offset)

    def lin(self, counts, physical=True):
        """Linear inversion (+ projection) -> device float64 [B, d, d, 2] (k3-k5)."""
        torch = nt.torch_cuda()
        if self.L_dev is None:
            raise nt.NativeError("plan was built without the linear-inversion table")
        counts = counts.reshape(-1, self.K).contiguous()
        B = counts.shape[0]
        rho = torch.empty((B, self.d, self.d, 2), dtype=torch.float64, device="cuda")
        nt.check(self._lib.qpb_lin_project(self.handle, B, nt.ptr(counts), int(bool(physical)), nt.ptr(rho),
                                           nt.stream_ptr()))
        return rho

    def lin_ordered(self, counts):
        """Physical linear estimates plus the start order for `mle` (qpb_lin_project_ordered)."""
        torch = nt.torch_cuda()
        if self.L_dev is None:
            raise nt.NativeError("plan was built without the linear-inversion table")
        counts = counts.reshape(-1, self.K).contiguous()
        B = counts.shape[0]
        rho = torch.empty((B, self.d, self.d, 2), dtype=torch.float64, device="cuda")
        order = torch.empty((B,), dtype=torch.int32, device="cuda")
        nt.check(self._lib.qpb_lin_project_ordered(self.handle, B, nt.ptr(counts), nt.ptr(rho), nt.ptr(order),
                                                   nt.stream_ptr()))
        return rho, order

    def mle(self, counts, rho0=None, max_iter=100, tol=1e-3, order=None):
        """R.rho.R maximum likelihood -> (rho [B, d, d, 2], iters [B] int32) on the device.  `order` (from
        `lin_ordered`) is a scheduling hint only."""
        torch = nt.torch_cuda()
        counts = counts.reshape(-1, self.K).contiguous()
        B = counts.shape[0]
        rho = torch.empty((B, self.d, self.d, 2), dtype=torch.float64, device="cuda")
        iters = torch.empty((B,), dtype=torch.int32, device="cuda")
        nt.check(self._lib.qpb_mle_rrr_ordered(self.handle, B, nt.ptr(counts), nt.ptr(rho0), nt.ptr(order), int(max_iter),
                                               float(tol), nt.ptr(rho), nt.ptr(iters), nt.stream_ptr()))
        return rho, iters

    def estimate(self, counts, method="lin", physical=True, init="lin", max_iter=100, tol=1e-3):
        """point_estimate dispatcher (state.py:143-189) on device counts -> (rho, iters or None)."""
        if method == "lin":
            return self.lin(counts, physical), None
        if method == "mle":
            order = None
            if init == "lin":
                start, order = self.lin_ordered(counts)  # state.py:209 -> point_estimate("lin"), physical by default
            elif init == "mixed":
                start = None
            else:
                raise ValueError("Invalid value for argument `init`")
            return self.mle(counts, start, max_iter, tol, order)
        raise ValueError("Invalid value for argument `method`")

    def bootstrap_buffers(self, n_samples, keep=False):
        """Device buffers for `bootstrap_into` (allocate once, reuse across calls)."""
        torch = nt.torch_cuda()
        B = int(n_samples)
        nbytes = self._lib.qpb_bootstrap_state_workspace(self.handle, B, self.P, self.O)
        return {
            "dist": torch.empty((B,), dtype=torch.float64, device="cuda"),
            "counts": torch.empty((B, self.P, self.O), dtype=torch.int32, device="cuda"),
            "iters": torch.empty((B,), dtype=torch.int32, device="cuda"),
            "rho": torch.empty((B, self.d, self.d, 2), dtype=torch.float64, device="cuda") if keep else None,
            "work": torch.empty((max(int(nbytes), 8),), dtype=torch.uint8, device="cuda"),
        }

    def bootstrap_into(self, bufs, probs, ref, seed, offset, method="lin", physical=True, init="lin",
                       max_iter=100, tol=1e-3, dst="hs"):
        """qpb_bootstrap_state on device-resident inputs (probs [K], ref [d, d, 2]) into `bufs`.
        Asynchronous: nothing is copied to or from the host."""
        if method not in nt.METHODS:
            raise ValueError("Invalid value for argument `method`")
        if init not in nt.INITS:
            raise ValueError("Invalid value for argument `init`")
        B = bufs["dist"].shape[0]
        shots = np.ascontiguousarray(self.n_shots)
        nt.check(self._lib.qpb_bootstrap_state(
            self.handle, B, self.P, self.O, nt.ptr(probs), shots.ctypes.data_as(ctypes.c_void_p),
            ctypes.c_uint64(seed), ctypes.c_uint64(offset), nt.METHODS[method], int(bool(physical)),
            nt.INITS[init], int(max_iter), float(tol), nt.ptr(ref), nt.DIST_KINDS[dst], nt.ptr(bufs["dist"]),
            nt.ptr(bufs["rho"]), nt.ptr(bufs["counts"]), nt.ptr(bufs["iters"]), nt.ptr(bufs["work"]),
            nt.stream_ptr()))
        return bufs

    def bootstrap(self, probs, n_samples, seed, offset, ref_matrix, method="lin", physical=True, init="lin",
                  max_iter=100, tol=1e-3, dst="hs", keep=False):
        """Fused bootstrap (interval.py:598-609) -> dict of device tensors: dist, counts, iters (+ rho)."""
        bufs = self.bootstrap_buffers(n_samples, keep)
        ref = nt.complex_to_device(ref_matrix)
        self.bootstrap_into(bufs, probs.contiguous(), ref, seed, offset, method, physical, init, max_iter, tol, dst)
        bufs["_keepalive"] = ref
        return bufs


# BootstrapStateInterval on one GPU goes through qpb_bootstrap_state_interval (tests switch it off to compare with the
# step-by-step path; both give the same bits)
FUSED_INTERVAL = True


def bootstrap_interval(plan, bloch, ref_matrix, n_samples, seed, offset, method, physical, init, max_iter, tol, dst,
                       levels):
    """qpb_bootstrap_state_interval: the whole single-GPU interval (interval.py:583-612) in one call with host
    inputs and outputs -> (quantiles at `levels` [host], sorted distances [device], iteration counts [device])."""
    torch = nt.torch_cuda()
    if method not in nt.METHODS:
        raise ValueError("Invalid value for argument `method`")
    if init not in nt.INITS:
        raise ValueError("Invalid value for argument `init`")
    B = int(n_samples)
    bloch = np.ascontiguousarray(bloch, dtype=np.float64).reshape(-1)
    ref = np.ascontiguousarray(ref_matrix, dtype=np.complex128)
    levels = np.ascontiguousarray(levels, dtype=np.float64).reshape(-1)
    if bloch.size != plan.D or ref.size != plan.D:
        raise ValueError("centre state does not match the POVM dimension")
    out = np.empty(levels.size, dtype=np.float64)
    dist = torch.empty((B,), dtype=torch.float64, device="cuda")
    iters = torch.empty((B,), dtype=torch.int32, device="cuda")
    # plain integers: the declared argtypes convert them (no ctypes objects built per call)
    rc = plan._lib.qpb_bootstrap_state_interval(
        plan.handle, B, plan.P, plan.O, plan.M_dev.data_ptr(), bloch.ctypes.data, ref.ctypes.data,
        plan.n_shots.ctypes.data, int(seed), int(offset), nt.METHODS[method],
        int(bool(physical)), nt.INITS[init], int(max_iter), float(tol), nt.DIST_KINDS[dst], levels.size,
        levels.ctypes.data, out.ctypes.data, dist.data_ptr(), iters.data_ptr(), nt.stream_ptr())
    if rc:
        nt.check(rc)
    return out, dist, iters


def quantiles_host(sorted_dev, levels):
    """qpb_quantiles_host: interp1d(linspace(0, 1, N), sorted)(levels) on a device-resident sorted array."""
    levels = np.ascontiguousarray(levels, dtype=np.float64)
    out = np.empty(levels.shape, dtype=np.float64)
    vp = ctypes.c_void_p
    nt.check(nt.load_library().qpb_quantiles_host(int(sorted_dev.numel()), nt.ptr(sorted_dev), levels.size,
                                                  vp(levels.ctypes.data), vp(out.ctypes.data), nt.stream_ptr()))
    return out


def sort_f64(values):
    """Ascending sort of a 1-D float64 device tensor (keys only) -> new device tensor."""
    torch = nt.torch_cuda()
    lib = nt.load_library()
    v = values.contiguous()
    out = torch.empty_like(v)
    nt.check(lib.qpb_sort_f64(int(v.numel()), nt.ptr(v), nt.ptr(out), nt.stream_ptr()))
    return out


def cholesky_vector(matrix):
    """Packed Cholesky parametrisation of a positive-definite matrix (quantpy/routines.py:84-91): the diagonal of
    the lower factor, then the real and the imaginary parts of its strict lower triangle (np.tril_indices order).
    Raises numpy.linalg.LinAlgError for a singular matrix, as the reference's la.cholesky does."""
    low = np.linalg.cholesky(np.asarray(matrix, dtype=np.complex128))
    strict = low[np.tril_indices(low.shape[0], -1)]
    return np.concatenate([np.real(np.diag(low)), strict.real, strict.imag])


def mhmc_chains(plan, counts, x_init, n_samples, step, burn_steps, thinning, deltas=None, uniforms=None, seed=0,
                chain_offset=0):
    """Metropolis-Hastings chains on the likelihood of `counts` (quantpy/mhmc.py:48-119 with normalized_update).
    x_init [C, D] packed Cholesky start vectors; counts [K] (shared) or [C, K].  deltas [C, T, D] / uniforms [C, T]
    (T = burn_steps + n_samples*thinning) replay a given noise stream; without them the kernel draws Philox noise.
    Returns dict(samples [C, n_samples, d, d, 2] device, accepted [C] device, x_final [C, D] device)."""
    torch = nt.torch_cuda()
    lib = nt.load_library()
    x0 = nt.to_device(np.asarray(x_init, dtype=np.float64).reshape(-1, plan.D), torch.float64)
    C = x0.shape[0]
    if torch.is_tensor(counts):
        cnt = counts.to(torch.int32).contiguous()
    else:
        cnt = nt.to_device(np.asarray(counts).reshape(-1, plan.K), torch.int32)
    batched = int(cnt.numel() == C * plan.K and C > 1)
    if not batched and cnt.numel() != plan.K:
        raise ValueError("counts must hold one table or one table per chain")
    total = int(burn_steps) + int(n_samples) * int(thinning)
    dl = ul = None
    if deltas is not None:
        dl = nt.to_device(np.asarray(deltas, dtype=np.float64).reshape(C, total, plan.D), torch.float64)
        ul = nt.to_device(np.asarray(uniforms, dtype=np.float64).reshape(C, total), torch.float64)
    samples = torch.empty((C, n_samples, plan.d, plan.d, 2), dtype=torch.float64, device="cuda")
    accepted = torch.zeros((C,), dtype=torch.int32, device="cuda")
    x_final = torch.empty((C, plan.D), dtype=torch.float64, device="cuda")
    nt.check(lib.qpb_mhmc_state(plan.handle, C, int(n_samples), int(thinning), int(burn_steps), float(step), nt.ptr(cnt),
                                batched, nt.ptr(x0), nt.ptr(dl), nt.ptr(ul), int(seed), int(chain_offset),
                                nt.ptr(samples), nt.ptr(accepted), nt.ptr(x_final), nt.stream_ptr()))
    return {"samples": samples, "accepted": accepted, "x_final": x_final}


def distance(rho, ref_matrix, dst="hs"):
    """dst(rho[b], ref) for a device batch rho [B, s, s, 2] -> device [B] (k8)."""
    torch = nt.torch_cuda()
    lib = nt.load_library()
    B, s = rho.shape[0], rho.shape[1]
    ref = nt.complex_to_device(ref_matrix)
    out = torch.empty((B,), dtype=torch.float64, device="cuda")
    nt.check(lib.qpb_distance(s, B, nt.ptr(rho.contiguous()), nt.ptr(ref), nt.DIST_KINDS[dst], nt.ptr(out),
                              nt.stream_ptr()))
    return out


class ProcessPlan:
    """Device tables for 'lifp' process estimation: wraps qpb_process_plan."""

    def __init__(self, n_qubits, lifp_oper_inv, S, K):
        nt.torch_cuda()
        lib = nt.load_library()
        self.n_qubits, self.S, self.K = n_qubits, S, K
        self.s = 4**n_qubits
        self.Linv_dev = nt.complex_to_device(np.asarray(lifp_oper_inv, dtype=np.complex128))
        handle = ctypes.c_void_p()
        nt.check(lib.qpb_process_plan_create(ctypes.byref(handle), n_qubits, S, K, nt.ptr(self.Linv_dev),
                                             nt.stream_ptr()))
        self.handle = handle
        self._lib = lib

    def __del__(self):
        handle = getattr(self, "handle", None)
        if handle:
            try:
                self._lib.qpb_process_plan_destroy(handle)
            except Exception:
                pass
            self.handle = None

    def lifp(self, counts, cptp=True, n_iter=1000, tol=1e-12):
        """counts device int32 [B, S, K] -> (choi [B, s, s, 2], iters [B])."""
        torch = nt.torch_cuda()
        counts = counts.reshape(-1, self.S * self.K).contiguous()
        B = counts.shape[0]
        choi = torch.empty((B, self.s, self.s, 2), dtype=torch.float64, device="cuda")
        iters = torch.empty((B,), dtype=torch.int32, device="cuda")
        nt.check(self._lib.qpb_lifp_cptp(self.handle, B, nt.ptr(counts), int(bool(cptp)), int(n_iter), float(tol),
                                         nt.ptr(choi), nt.ptr(iters), nt.stream_ptr()))
        return choi, iters


def cptp_project(choi_matrices, n_qubits, n_iter=1000, tol=1e-12):
    """Alternating TP/CP projection (process.py:231-257) of host Choi matrices [B, s, s] on the GPU."""
    torch = nt.torch_cuda()
    lib = nt.load_library()
    x = nt.complex_to_device(np.asarray(choi_matrices, dtype=np.complex128))
    B = x.shape[0]
    out = torch.empty_like(x)
    iters = torch.empty((B,), dtype=torch.int32, device="cuda")
    nt.check(lib.qpb_cptp_project(n_qubits, B, nt.ptr(x), int(n_iter), float(tol), nt.ptr(out), nt.ptr(iters),
                                  nt.stream_ptr()))
    return out, iters


def choi_from_states(G, rho, n_qubits):
    """choi[b] = sum_s G_s (x) rho[b, s]: G host complex [S, d, d]; rho device [B, S, d, d, 2] -> device [B, d^2, d^2, 2]."""
    torch = nt.torch_cuda()
    lib = nt.load_library()
    B, S = rho.shape[0], rho.shape[1]
    s = 4**n_qubits
    Gd = nt.complex_to_device(np.asarray(G, dtype=np.complex128))
    choi = torch.empty((B, s, s, 2), dtype=torch.float64, device="cuda")
    nt.check(lib.qpb_choi_from_states(n_qubits, S, B, nt.ptr(Gd), nt.ptr(rho.contiguous()), nt.ptr(choi),
                                      nt.stream_ptr()))
    return choi


def cptp_project_if_needed(choi, n_qubits, n_iter=1000, tol=1e-12, atol=1e-5):
    """Device Choi batch -> (projected batch, iters); matrices passing Channel.is_cptp(atol) are left untouched."""
    torch = nt.torch_cuda()
    lib = nt.load_library()
    B = choi.shape[0]
    out = torch.empty_like(choi)
    iters = torch.empty((B,), dtype=torch.int32, device="cuda")
    nt.check(lib.qpb_cptp_project_if_needed(n_qubits, B, nt.ptr(choi.contiguous()), int(n_iter), float(tol),
                                            float(atol), nt.ptr(out), nt.ptr(iters), nt.stream_ptr()))
    return out, iters


def l2_moments(frequencies, n_trials, weights):
    """(mean, variance) of the weighted squared l2 error (quantpy/stats.py:5-53) on the GPU.
    frequencies [P, O] or [B, P, O]; weights [P, O, P, O] -> floats or arrays of length B."""
    torch = nt.torch_cuda()
    lib = nt.load_library()
    f = np.asarray(frequencies, dtype=np.float64)
    single = f.ndim == 2
    f3 = f[None] if single else f
    B, P, O = f3.shape
    w = nt.to_device(np.asarray(weights, dtype=np.float64).reshape(P, O, P, O), torch.float64)
    fd = nt.to_device(f3, torch.float64)
    mean = torch.empty((B,), dtype=torch.float64, device="cuda")
    var = torch.empty((B,), dtype=torch.float64, device="cuda")
    nt.check(lib.qpb_l2_moments(B, P, O, nt.ptr(w), nt.ptr(fd), float(n_trials), nt.ptr(mean), nt.ptr(var),
                                nt.stream_ptr()))
    m, v = mean.cpu().numpy(), var.cpu().numpy()
    return (float(m[0]), float(v[0])) if single else (m, v)


_PLAN_CACHE = {}      # content digest -> StatePlan
_PLAN_BY_IDENTITY = {}  # (id(povm array), id(shots array)) -> (weak refs, cheap fingerprints, plan)


def _fingerprint(a):
    """Cheap change detector for an array whose identity we have seen: shape, strides, dtype, data pointer and
    a strided sample of at most 64 values (a full hash of the 2.65 MB POVM tensor at n = 4 cost more than the
    kernels of a 1000-sample interval)."""
    flat = a.reshape(-1)
    step = max(1, flat.size // 64)
    return (a.shape, a.dtype.str, a.__array_interface__["data"][0], flat[::step][:64].tobytes())


def state_plan(povm_matrix, n_measurements):
    """Plans are cached so that repeated point_estimate / setup calls reuse the uploaded tables.

    Fast path: the same array OBJECTS as last time (the tomograph's `povm_matrix` / `n_measurements` attributes)
    with unchanged fingerprints -> no hashing.  Slow path: content digest over both arrays (arrays rebuilt by
    generate_measurement_matrix on every call still hit the cache)."""
    import weakref

    ident = None
    if isinstance(povm_matrix, np.ndarray) and isinstance(n_measurements, np.ndarray):
        ident = (id(povm_matrix), id(n_measurements))
        hit = _PLAN_BY_IDENTITY.get(ident)
        if hit is not None:
            ref_p, ref_n, fp_p, fp_n, plan = hit
            if ref_p() is povm_matrix and ref_n() is n_measurements and _fingerprint(povm_matrix) == fp_p \
                    and _fingerprint(n_measurements) == fp_n:
                return plan
    pm = np.ascontiguousarray(np.asarray(povm_matrix, dtype=np.float64))
    n = np.ascontiguousarray(np.asarray(n_measurements, dtype=np.float64).reshape(-1))
    key = (pm.shape, hashlib.blake2b(pm.tobytes() + n.tobytes(), digest_size=16).digest())
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        if len(_PLAN_CACHE) >= 16:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
        plan = _PLAN_CACHE[key] = StatePlan(pm, n)
    if ident is not None:
        if len(_PLAN_BY_IDENTITY) >= 64:
            _PLAN_BY_IDENTITY.clear()
        try:
            _PLAN_BY_IDENTITY[ident] = (weakref.ref(povm_matrix), weakref.ref(n_measurements),
                                        _fingerprint(povm_matrix), _fingerprint(n_measurements), plan)
        except TypeError:  # pragma: no cover  (array subclass without weakref support)
            pass
    return plan
