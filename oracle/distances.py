"""Distance oracle (test infrastructure only).  Follows quantpy/geometry.py:5-56.

Two forms are given for each SciPy-``sqrtm`` based distance: ``*_sqrtm`` is the
reference's literal expression; the plain name is the eigenvalue form the CUDA
kernel implements.  They agree to ~1e-14 on Hermitian inputs whose spectrum is
not within ~1e-8 of singular (``sqrtm`` itself loses accuracy there); the golden
tests state the tolerance used.
All functions reproduce the reference's "below 1e-15 -> integer 0" rule.
"""

import numpy as np
import scipy.linalg as la

ZERO_BELOW = 1e-15


def _snap(x):
    x = np.asarray(x, dtype=float)
    return np.where(x < ZERO_BELOW, 0.0, x)


def hs(a, b):
    """sqrt|Tr (a-b)^2| / sqrt 2.  geometry.py:16.  Batched over leading axes."""
    delta = np.asarray(a) - np.asarray(b)
    tr = np.einsum("...ij,...ji->...", delta, delta)
    return _snap(np.sqrt(np.abs(tr)) / np.sqrt(2))


def trace(a, b):
    """|Tr sqrt((a-b)^2)| / 2 = sum |eig(a-b)| / 2 for Hermitian a-b.  geometry.py:34."""
    delta = np.asarray(a) - np.asarray(b)
    delta = 0.5 * (delta + np.conj(np.swapaxes(delta, -1, -2)))
    return _snap(np.sum(np.abs(np.linalg.eigvalsh(delta)), axis=-1) / 2)


def trace_sqrtm(a, b):
    delta = np.asarray(a) - np.asarray(b)
    return float(_snap(abs(np.trace(la.sqrtm(delta @ delta))) / 2))


def _psd_sqrt(m):
    vals, vecs = np.linalg.eigh(m)
    vals = np.sqrt(np.maximum(vals, 0.0))
    return (vecs * vals[..., None, :]) @ np.conj(np.swapaxes(vecs, -1, -2))


def infidelity(a, b):
    """1 - |Tr sqrt(sqrt(a) b sqrt(a))|^2.  geometry.py:52, eigenvalue form."""
    a = np.asarray(a)
    b = np.asarray(b)
    ra = _psd_sqrt(0.5 * (a + np.conj(np.swapaxes(a, -1, -2))))
    inner = ra @ b @ ra
    inner = 0.5 * (inner + np.conj(np.swapaxes(inner, -1, -2)))
    vals = np.maximum(np.linalg.eigvalsh(inner), 0.0)
    return _snap(1 - np.sum(np.sqrt(vals), axis=-1) ** 2)


def infidelity_sqrtm(a, b):
    ra = la.sqrtm(a)
    return float(_snap(1 - np.abs(np.trace(la.sqrtm(ra @ b @ ra)) ** 2)))


BY_NAME = {"hs": hs, "trace": trace, "if": infidelity}
