"""Pauli basis and Bloch<->matrix maps (oracle; test infrastructure only).

Follows quantpy/routines.py:5-19 (Pauli order I,X,Y,Z; n-qubit index
i1*4^(n-1)+...+in, first qubit most significant) and quantpy/qobj.py:109-135
(matrix = sum_i bloch_i sigma_i ; bloch_i = Re Tr(sigma_i rho^dagger) / 2^n).
"""

from functools import lru_cache

import numpy as np

_S1 = np.zeros((4, 2, 2), dtype=np.complex128)
_S1[0] = [[1, 0], [0, 1]]
_S1[1] = [[0, 1], [1, 0]]
_S1[2] = [[0, -1j], [1j, 0]]
_S1[3] = [[1, 0], [0, -1]]


@lru_cache(maxsize=None)
def pauli_basis(n_qubits):
    """(4^n, 2^n, 2^n) complex array of Pauli strings, reference ordering."""
    out = _S1
    for _ in range(n_qubits - 1):
        a, b = out.shape[0], out.shape[1]
        out = np.einsum("iab,jcd->ijacbd", out, _S1).reshape(a * 4, b * 2, b * 2)
    out = np.ascontiguousarray(out)
    out.setflags(write=False)
    return out


def n_qubits_from_D(D):
    n = int(round(np.log2(D) / 2))
    if 4**n != D:
        raise ValueError("length is not a power of 4")
    return n


def bloch_to_matrix(bloch):
    """(..., 4^n) real -> (..., 2^n, 2^n) complex.  qobj.py:109-118."""
    bloch = np.asarray(bloch)
    n = n_qubits_from_D(bloch.shape[-1])
    return np.tensordot(bloch, pauli_basis(n), axes=([-1], [0]))


def matrix_to_bloch(matrix):
    """(..., d, d) complex -> (..., d^2) real.  qobj.py:126-135 with
    geometry.py:59-70: Re Tr(sigma_i @ conj(rho.T)) / d."""
    matrix = np.asarray(matrix)
    d = matrix.shape[-1]
    n = int(round(np.log2(d)))
    S = pauli_basis(n)
    # Tr(S_i @ conj(M^T)) = sum_ab S_i[a,b] * conj(M[a,b])
    return np.real(np.einsum("iab,...ab->...i", S, np.conj(matrix))) / d
