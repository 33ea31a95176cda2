"""Distances between quantum objects (host mirror of quantpy/geometry.py).

These are the single-pair, host-side forms used by user code and by custom `dst` callables.
The bootstrap loop does NOT call them: it computes the same quantities for a whole batch in
csrc/state.cu:k_distance (eigenvalue forms).  Each returns the integer 0 below 1e-15 like the reference.
"""

import numpy as np
import scipy.linalg as la

_ZERO_BELOW = 1e-15


def _as_matrix(obj):
    return obj if isinstance(obj, np.ndarray) else obj.matrix


def _snap(value):
    return 0 if value < _ZERO_BELOW else value


def hs_dst(A, B):
    """Hilbert-Schmidt distance sqrt|Tr (A-B)^2| / sqrt 2 (quantpy/geometry.py:5-20)."""
    delta = _as_matrix(A) - _as_matrix(B)
    return _snap(np.sqrt(abs(np.einsum("ij,ji->", delta, delta))) / np.sqrt(2))


def trace_dst(A, B):
    """Trace distance |Tr sqrt((A-B)^2)| / 2 (quantpy/geometry.py:23-38)."""
    delta = _as_matrix(A) - _as_matrix(B)
    return _snap(abs(np.trace(la.sqrtm(delta @ delta))) / 2)


def if_dst(A, B):
    """Infidelity 1 - |Tr sqrt(sqrt(A) B sqrt(A))|^2 (quantpy/geometry.py:41-56)."""
    root = la.sqrtm(_as_matrix(A))
    return _snap(1 - np.abs(np.trace(la.sqrtm(root @ _as_matrix(B) @ root)) ** 2))


def product(A, B):
    """Hermitian inner product Tr(A B^dagger) (quantpy/geometry.py:59-70)."""
    a, b = _as_matrix(A), _as_matrix(B)
    return np.sum(a * np.conj(b), dtype=np.complex128)
