"""Process-tomography oracle (test infrastructure only; never imported by quantpy_b200).

NumPy restatement of quantpy/tomography/process.py:

* input states ......... process.py:330-339
* experiment ........... process.py:91-129 (one state tomography per input state)
* 'lifp' operator ...... process.py:194-211, estimate process.py:284-289
* CPTP projection ...... process.py:231-278 (alternating TP / CP projections)
* 'states' assembly .... process.py:316-327, quantpy/basis.py:7-42
* channel application .. quantpy/channel.py:131-142, Choi build channel.py:95-103

Vectorisation is column stacking (routines.py:53-61): vec(M)[c*s + r] = M[r, c].
A channel is represented by its Choi matrix sum_ij E_ij (x) Phi(E_ij) (input factor first).
"""

import numpy as np
import scipy.linalg as la

from . import state as ostate
from .pauli import bloch_to_matrix, n_qubits_from_D

CP_CLIP = 1e-12  # EPS in process.py:272


def mat2vec(m):
    return np.swapaxes(m, -1, -2).reshape(m.shape[:-2] + (-1,))


def vec2mat(v):
    s = int(round(np.sqrt(v.shape[-1])))
    return np.swapaxes(v.reshape(v.shape[:-1] + (s, s)), -1, -2)


# ----------------------------------------------------------------------------- channels

def choi_from_map(fn, n_qubits):
    """Choi matrix of a linear map fn: (d,d)->(d,d).  channel.py:95-103."""
    d = 2**n_qubits
    out = np.zeros((d * d, d * d), dtype=np.complex128)
    for i in range(d):
        for j in range(d):
            unit = np.zeros((d, d), dtype=np.complex128)
            unit[i, j] = 1
            out += np.kron(unit, fn(unit))
    return out


def depolarizing_choi(p, n_qubits):
    """rho -> p Tr(rho) I/d + (1-p) rho.  channel.py:232-236."""
    d = 2**n_qubits
    return choi_from_map(lambda x: p * np.trace(x) * np.eye(d) / d + (1 - p) * x, n_qubits)


def apply_choi(choi, rho):
    """Tr_in[(rho^T (x) I) C].  channel.py:139-141."""
    d = rho.shape[-1]
    c4 = np.asarray(choi).reshape(d, d, d, d)  # [i, a, j, b]
    return np.einsum("ij,iajb->ab", np.asarray(rho), c4)


def input_states(spec, n_qubits):
    """process.py:330-339: Bloch rows of the named POVM, each normalised to unit trace."""
    if isinstance(spec, (list, tuple)):
        return [np.asarray(s, dtype=np.complex128) for s in spec]
    rows = np.squeeze(ostate.measurement_matrix(spec, n_qubits))
    mats = bloch_to_matrix(rows)
    return [m / np.trace(m) for m in mats]


# ----------------------------------------------------------------------------- experiment

def experiment(choi, inputs, povm, n_meas, size=None, rng=None):
    """Counts (S, P, O) [or (size, S, P, O)].  process.py:121-129."""
    from .pauli import matrix_to_bloch

    per_state = []
    for rho_in in inputs:
        out = apply_choi(choi, rho_in)
        per_state.append(ostate.experiment(povm, matrix_to_bloch(out), n_meas, size=size, rng=rng))
    return np.stack(per_state, axis=-3)


# ----------------------------------------------------------------------------- lifp

def lifp_operator(inputs, povm, n_meas):
    """(S*K, d^4) complex operator whose rows are vec(rho_in (x) E_k^T).  process.py:197-208."""
    A = ostate.weighted_povm(povm, n_meas)
    E = bloch_to_matrix(A)  # (K, d, d)
    rows = [mat2vec(np.kron(rho_in, e.T)) for rho_in in inputs for e in E]
    return np.array(rows)


def lifp_estimate(counts, inputs, povm, n_meas=None, cptp=True, n_iter=1000, tol=1e-12,
                  return_iters=False):
    """Linear-inversion Choi estimate (+ optional CPTP projection).  process.py:284-289.

    counts: (S, P, O) or (B, S, P, O).  Frequencies are normalised PER INPUT STATE
    (process.py:285); the left inverse is the plain-transpose one (routines.py:69-71).
    """
    counts = np.asarray(counts)
    batched = counts.ndim == 4
    c = counts if batched else counts[None]
    if n_meas is None:
        n_meas = c[0, 0].sum(-1)
    B, S = c.shape[:2]
    flat = c.reshape(B, S, -1).astype(float)
    freq = (flat / flat.sum(-1, keepdims=True)).reshape(B, -1)
    linv = ostate.left_inverse(lifp_operator(inputs, povm, n_meas))
    choi = vec2mat(freq @ linv.T)
    iters = np.zeros(B, dtype=np.int32)
    if cptp:
        choi, iters = cptp_projection(choi, n_iter=n_iter, tol=tol, return_iters=True)
    out = choi if batched else choi[0]
    if return_iters:
        return out, (iters if batched else int(iters[0]))
    return out


# ----------------------------------------------------------------------------- CPTP projection

def ptrace_operator(n_qubits):
    """(d^2, d^4) operator taking vec(Choi) to vec(Tr_out Choi).  routines.py:47-50."""
    d = 2**n_qubits
    eye = np.eye(d)
    total = 0
    for unit in eye:
        total = total + np.kron(eye, np.kron(unit, np.kron(eye, unit)))
    return total


def tp_projection_vec(vec, P, PtP, d):
    """process.py:259-268 on vectorised Choi matrices (..., d^4)."""
    shift = P.T.conj() @ mat2vec(np.eye(d))
    return vec + (shift - vec @ PtP.T) / d


def cp_projection(choi):
    """process.py:270-278: eigh, clip at 1e-12, recompose (no renormalisation)."""
    vals, vecs = np.linalg.eigh(choi)
    vals = np.maximum(CP_CLIP, vals)
    return (vecs * vals[..., None, :]) @ np.conj(np.swapaxes(vecs, -1, -2))


def cptp_projection(choi, n_iter=1000, tol=1e-12, return_iters=False):
    """Alternating-projection loop of process.py:237-257, batched over leading axis.

        y' = TP(x + p);  x' = CP(y' + q)
        crit = 2(|<y'-y, q>| + |<x'-x, p>|)          (old p, q)
        p += x' - y';  q += y' - x'
        crit += ||x'-y'||^2 + ||y'-x'||^2 ;  stop when crit < tol
    """
    choi = np.asarray(choi, dtype=np.complex128)
    single = choi.ndim == 2
    x = mat2vec(choi if not single else choi[None]).copy()
    B, L = x.shape
    d = int(round(L ** 0.25))
    n = int(round(np.log2(d)))
    P = ptrace_operator(n)
    PtP = P.T.conj() @ P
    p = np.zeros_like(x)
    q = np.zeros_like(x)
    y = np.zeros_like(x)
    iters = np.zeros(B, dtype=np.int32)
    active = np.arange(B)
    for it in range(1, int(n_iter) + 1):
        if active.size == 0:
            break
        xa, pa, qa, ya = x[active], p[active], q[active], y[active]
        y_new = tp_projection_vec(xa + pa, P, PtP, d)
        x_new = mat2vec(cp_projection(vec2mat(y_new + qa)))
        y_diff = y_new - ya
        x_diff = x_new - xa
        crit = 2 * (np.abs(np.sum(np.conj(y_diff) * qa, -1)) + np.abs(np.sum(np.conj(x_diff) * pa, -1)))
        p_diff = x_new - y_new
        crit = crit + 2 * np.sum(np.abs(p_diff) ** 2, -1)
        x[active], y[active] = x_new, y_new
        p[active], q[active] = pa + p_diff, qa - p_diff
        iters[active] = it
        active = active[~(crit < tol)]
    out = vec2mat(x)
    out = out[0] if single else out
    if return_iters:
        return out, (int(iters[0]) if single else iters)
    return out


# ----------------------------------------------------------------------------- 'states' method

def _gram(elements):
    """basis.py:20-29 with the trace product geometry.py:59-70."""
    e = np.asarray(elements)
    return np.einsum("iab,jab->ij", e, np.conj(e))


def basis_decompose(elements, obj):
    """basis.py:31-34."""
    e = np.asarray(elements)
    rhs = np.einsum("iab,ab->i", e, np.conj(obj))
    return np.conj(la.solve(_gram(e), rhs))


def states_estimate(counts, inputs, povm, n_meas=None, cptp=True, method="lin", physical=True,
                    init="lin", max_iter=1000, tol=1e-10, mle="bfgs"):
    """process.py:316-327: reconstruct each output state, then assemble the Choi matrix
    through the input basis.  ``mle`` selects which MLE oracle stands in for
    StateTomograph.point_estimate('mle') ('bfgs' = reference algorithm, 'rrr' = kernel spec).
    Note the reference forwards (n_iter, tol) positionally as (max_iter, tol), process.py:317."""
    counts = np.asarray(counts)
    inputs = [np.asarray(s) for s in inputs]
    d = inputs[0].shape[0]
    outs = []
    for c in counts:
        if method == "lin":
            outs.append(ostate.lin_estimate(c, povm, n_meas, physical=physical))
        elif method == "mle":
            fn = ostate.mle_bfgs if mle == "bfgs" else ostate.mle_rrr
            outs.append(fn(c, povm, n_meas, init=init, max_iter=max_iter, tol=tol))
        else:
            raise ValueError("Invalid value for argument `method`")
    outs = np.asarray(outs)
    choi = np.zeros((d * d, d * d), dtype=np.complex128)
    for i in range(d):
        for j in range(d):
            unit = np.zeros((d, d))
            unit[i, j] = 1
            coef = basis_decompose(inputs, unit)
            choi += np.kron(np.tensordot(coef, np.asarray(inputs), 1), np.tensordot(coef, outs, 1))
    if cptp and not is_cptp(choi):
        choi = cptp_projection(choi)
    return choi


def is_cptp(choi, atol=1e-5):
    """channel.py:144-157 (non-Hermitian eigenvalues via la.eig like qobj.py:202)."""
    dd = choi.shape[0]
    d = int(round(np.sqrt(dd)))
    rho_in = np.einsum("iaja->ij", choi.reshape(d, d, d, d))
    tp = np.allclose(rho_in, np.eye(d), atol=atol)
    cp = np.allclose(np.minimum(np.real(la.eig(choi)[0]), 0), 0, atol=atol)
    return bool(tp and cp)
