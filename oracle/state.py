"""State-tomography oracle (test infrastructure only; never imported by quantpy_b200).

NumPy/SciPy restatement of the state path of nordmtr/quantpy:

* POVM tables ............ quantpy/measurements.py:4-94
* probabilities/sampling . quantpy/tomography/state.py:99-128
* linear inversion ....... quantpy/tomography/state.py:191-202, quantpy/routines.py:69-71
* physical projection .... quantpy/tomography/state.py:267-273
* BFGS/Cholesky "mle" .... quantpy/tomography/state.py:204-229, quantpy/routines.py:84-101
* R.rho.R MLE ............ NOT in the reference; specification of the CUDA kernel
                           (likelihood = state.py:217-229 including its 1e-10 guard)

All functions take plain arrays.  ``povm`` is the reference's 3-D POVM tensor
(P, O, D) of Pauli coefficients; ``counts`` is (P, O) integers (or (B, P, O) for
the batched variants); ``n_meas`` is the (P,) shot vector.
"""

import numpy as np
import scipy.linalg as la
from scipy.optimize import minimize

from .pauli import bloch_to_matrix, matrix_to_bloch, n_qubits_from_D, pauli_basis

LOG_GUARD = 1e-10  # EPS in state.py:219
CLIP_EIG = 1e-15  # EPS in state.py:269

# ----------------------------------------------------------------------------- POVM tables

_AX = {"x+": (1, 1, 0, 0), "x-": (1, -1, 0, 0), "y+": (1, 0, 1, 0),
       "y-": (1, 0, -1, 0), "z+": (1, 0, 0, 1), "z-": (1, 0, 0, -1)}


def _one_qubit_povm(name):
    """Single-qubit proto-POVMs, measurements.py:36-72."""
    g = lambda *keys: np.array([_AX[k] for k in keys], dtype=float)  # noqa: E731
    if name == "proj":
        return g("x+", "x-", "y+", "y-", "z+", "z-")[None] / 6
    if name == "proj-set":
        return np.stack([g("x+", "x-"), g("y+", "y-"), g("z+", "z-")]) / 2
    if name == "proj4":
        return g("x+", "y+", "z+", "z-")[None] / 4
    if name == "sic":
        s = 1 / np.sqrt(3)
        return np.array([[1, s, s, s], [1, s, -s, -s], [1, -s, s, -s], [1, -s, -s, s]])[None] / 4
    raise ValueError("Incorrect string shortcut for argument `povm`")


def measurement_matrix(povm="proj", n_qubits=1):
    """POVM tensor (P, O, 4^n).  measurements.py:33-94."""
    if isinstance(povm, str):
        one = _one_qubit_povm(povm)
    else:
        povm = np.asarray(povm)
        if povm.shape[-1] == 4:  # single-qubit table, checked first (measurements.py:77)
            one = povm[None] if povm.ndim == 2 else povm
        elif povm.shape[-1] == 4**n_qubits:
            return povm[None] if povm.ndim == 2 else povm
        else:
            raise ValueError("Incorrect POVM matrix")
    full = one
    for _ in range(n_qubits - 1):
        full = np.kron(full, one)
    return full


def weighted_povm(povm, n_meas):
    """(K, D) matrix of shot-weighted POVM rows.  state.py:193-196 / 221-224."""
    povm = np.asarray(povm, dtype=float)
    n_meas = np.asarray(n_meas, dtype=float)
    return (povm * n_meas[:, None, None] / n_meas.sum()).reshape(-1, povm.shape[-1])


def probabilities(povm, bloch):
    """clip(2^n * M.r, 0, 1) per POVM.  state.py:109-110."""
    povm = np.asarray(povm, dtype=float)
    n = n_qubits_from_D(povm.shape[-1])
    return np.clip(np.einsum("pok,k->po", povm, np.asarray(bloch, dtype=float)) * 2**n, 0, 1)


def experiment(povm, bloch, n_meas, size=None, rng=None):
    """Shot sampling, one multinomial per POVM.  state.py:111-114.

    ``rng=None`` uses the global legacy ``np.random`` stream like the reference.
    Returns (P, O) counts, or (size, P, O) when ``size`` is given.
    """
    rng = np.random if rng is None else rng
    probs = probabilities(povm, bloch)
    cols = [rng.multinomial(int(n), p, size=size) for p, n in zip(probs, np.asarray(n_meas))]
    return np.stack(cols, axis=-2)


# ----------------------------------------------------------------------------- linear inversion

def left_inverse(A):
    """(A^T A)^-1 A^T with a PLAIN transpose.  routines.py:69-71."""
    return la.inv(A.T @ A) @ A.T


def make_feasible(rho):
    """Clip eigenvalues at 1e-15, recompose, renormalise.  state.py:267-273.
    Accepts (..., d, d)."""
    vals, vecs = np.linalg.eigh(rho)
    vals = np.maximum(CLIP_EIG, vals)
    out = (vecs * vals[..., None, :]) @ np.conj(np.swapaxes(vecs, -1, -2))
    return out / np.trace(out, axis1=-2, axis2=-1)[..., None, None]


def lin_estimate(counts, povm, n_meas=None, physical=True):
    """Linear-inversion estimate.  state.py:191-202.

    counts: (P, O) or (B, P, O).  Returns (d, d) or (B, d, d) complex.
    """
    povm = np.asarray(povm, dtype=float)
    counts = np.asarray(counts)
    batched = counts.ndim == 3
    c = counts if batched else counts[None]
    if n_meas is None:
        n_meas = c[0].sum(-1)  # results setter, state.py:138-141
    n = n_qubits_from_D(povm.shape[-1])
    flat = c.reshape(c.shape[0], -1)
    freq = flat / flat.sum(-1, keepdims=True)
    linv = left_inverse(weighted_povm(povm, n_meas))
    rho = bloch_to_matrix(freq @ linv.T / 2**n)
    if physical:
        rho = make_feasible(rho)
    return rho if batched else rho[0]


# ----------------------------------------------------------------------------- reference "mle" (BFGS over a Cholesky factor)

def _chol_pack(rho):
    """routines.py:84-91: diag (real part) then Re, Im of the strict lower triangle."""
    low = la.cholesky(rho, lower=True)
    d = rho.shape[0]
    strict = low[np.tril_indices(d, -1)]
    return np.concatenate([np.real(np.diag(low)), strict.real, strict.imag])


def _chol_unpack(x):
    """routines.py:94-101: T T^dagger from the packed lower-triangular factor."""
    d = int(np.sqrt(len(x)))
    rest = x[d:]
    half = len(rest) // 2
    low = np.zeros((d, d), dtype=np.complex128)
    low[np.tril_indices(d, -1)] = rest[:half] + 1j * rest[half:]
    low[np.diag_indices(d)] = x[:d]
    return low @ low.conj().T


def neg_log_likelihood(rho, counts, povm, n_meas):
    """-sum_k f_k log(p_k + 1e-10).  state.py:217-229 (p from the weighted POVM)."""
    n = n_qubits_from_D(np.asarray(povm).shape[-1])
    A = weighted_povm(povm, n_meas)
    p = A @ matrix_to_bloch(rho) * 2**n
    f = np.asarray(counts).reshape(-1) / np.sum(n_meas)
    return -np.sum(f * np.log(p + LOG_GUARD))


def mle_bfgs(counts, povm, n_meas=None, init="lin", max_iter=100, tol=1e-3):
    """The reference's 'mle': SciPy BFGS (finite-difference gradient) over the
    Cholesky parameters of an unnormalised rho.  state.py:204-215."""
    povm = np.asarray(povm, dtype=float)
    counts = np.asarray(counts)
    if n_meas is None:
        n_meas = counts.sum(-1)
    n = n_qubits_from_D(povm.shape[-1])
    if init == "mixed":
        start = np.eye(2**n, dtype=np.complex128) / 2**n
    elif init == "lin":
        start = lin_estimate(counts, povm, n_meas, physical=True)
    else:
        raise ValueError("Invalid value for argument `init`")
    A = weighted_povm(povm, n_meas)
    f = counts.reshape(-1) / np.sum(n_meas)
    S = pauli_basis(n)
    d = 2**n

    def nll(x):
        m = _chol_unpack(x)
        m = m / np.trace(m)
        bloch = np.real(np.einsum("iab,ab->i", S, np.conj(m))) / d
        return -np.sum(f * np.log(A @ bloch * d + LOG_GUARD))

    res = minimize(nll, _chol_pack(start), method="BFGS", tol=tol, options={"maxiter": max_iter})
    m = _chol_unpack(res.x)
    return m / np.trace(m)


def mle_slsqp(counts, povm, n_meas=None, init="lin", max_iter=100, tol=1e-3):
    """The reference's 'mle-constr': the same Cholesky-parametrised likelihood as `mle_bfgs`, minimised by
    SciPy SLSQP under the equality Tr(L L^dagger) = 1.  state.py:231-254, 262-265."""
    povm = np.asarray(povm, dtype=float)
    counts = np.asarray(counts)
    if n_meas is None:
        n_meas = counts.sum(-1)
    n = n_qubits_from_D(povm.shape[-1])
    if init == "mixed":
        start = np.eye(2**n, dtype=np.complex128) / 2**n
    elif init == "lin":
        start = lin_estimate(counts, povm, n_meas, physical=True)
    else:
        raise ValueError("Invalid value for argument `init`")
    A = weighted_povm(povm, n_meas)
    f = counts.reshape(-1) / np.sum(n_meas)
    S = pauli_basis(n)
    d = 2**n

    def nll(x):
        m = _chol_unpack(x)
        m = m / np.trace(m)
        bloch = np.real(np.einsum("iab,ab->i", S, np.conj(m))) / d
        return -np.sum(f * np.log(A @ bloch * d + LOG_GUARD))

    # the reference returns the complex trace; SLSQP keeps its real part (ComplexWarning there)
    unit_trace = [{"type": "eq", "fun": lambda x: np.real(np.trace(_chol_unpack(x))) - 1}]
    res = minimize(nll, _chol_pack(start), constraints=unit_trace, method="SLSQP", tol=tol,
                   options={"maxiter": max_iter})
    m = _chol_unpack(res.x)
    return m / np.trace(m)


# ----------------------------------------------------------------------------- R.rho.R MLE (kernel specification)

def povm_operators(A):
    """E_k = sum_i A[k,i] sigma_i as (K, d, d) complex."""
    return bloch_to_matrix(A)


def mle_rrr(counts, povm, n_meas=None, rho0=None, init="lin", max_iter=100, tol=1e-3,
            return_iters=False):
    """Iterative maximum likelihood  rho <- R rho R / Tr(R rho R).

    Specification of the CUDA kernel ``qpb_mle_rrr`` (no reference counterpart):

        f_k = counts_k / sum(counts)                     (state.py:227)
        A   = shot-weighted POVM rows, E_k = sum_i A_ki sigma_i   (state.py:221-224)
        repeat up to max_iter times:
            p_k  = Re Tr(E_k rho)                        (= 2^n A.bloch(rho), state.py:226)
            R    = sum_k f_k / (p_k + 1e-10) E_k         (gradient of state.py:228 incl. guard)
            rho' = R rho R ; rho' = (rho' + rho'^dagger)/2 ; rho' /= Tr rho'
            delta = ||rho' - rho||_F ; rho = rho'
            stop when delta < tol
    Start: ``rho0`` if given, else the physical 'lin' estimate (init='lin') or I/d
    (init='mixed') exactly as state.py:206-211 picks the BFGS start.

    counts: (P, O) or (B, P, O).  Returns rho (and the per-sample iteration count).
    """
    povm = np.asarray(povm, dtype=float)
    counts = np.asarray(counts)
    batched = counts.ndim == 3
    c = counts if batched else counts[None]
    B = c.shape[0]
    if n_meas is None:
        n_meas = c[0].sum(-1)
    n = n_qubits_from_D(povm.shape[-1])
    d = 2**n
    if rho0 is None:
        if init == "lin":
            rho = lin_estimate(c, povm, n_meas, physical=True)
        elif init == "mixed":
            rho = np.broadcast_to(np.eye(d, dtype=np.complex128) / d, (B, d, d)).copy()
        else:
            raise ValueError("Invalid value for argument `init`")
    else:
        rho = np.array(rho0, dtype=np.complex128).reshape(-1, d, d)
        if rho.shape[0] == 1 and B > 1:
            rho = np.repeat(rho, B, axis=0)
    rho = np.array(rho, dtype=np.complex128)
    E = povm_operators(weighted_povm(povm, n_meas))
    flat = c.reshape(B, -1).astype(float)
    f = flat / flat.sum(-1, keepdims=True)
    iters = np.zeros(B, dtype=np.int32)
    active = np.arange(B)
    for it in range(1, int(max_iter) + 1):
        if active.size == 0:
            break
        cur = rho[active]
        p = np.real(np.einsum("kab,nba->nk", E, cur))
        w = f[active] / (p + LOG_GUARD)
        R = np.einsum("nk,kab->nab", w, E)
        new = R @ cur @ R
        new = 0.5 * (new + np.conj(np.swapaxes(new, -1, -2)))
        new /= np.real(np.trace(new, axis1=-2, axis2=-1))[:, None, None]
        delta = np.sqrt(np.sum(np.abs(new - cur) ** 2, axis=(-2, -1)))
        rho[active] = new
        iters[active] = it
        active = active[~(delta < tol)]
    out = rho if batched else rho[0]
    if return_iters:
        return out, (iters if batched else int(iters[0]))
    return out
