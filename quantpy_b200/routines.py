"""Small host-side linear-algebra helpers shared by the quantum-object classes.

Host mirror of quantpy/routines.py (same public names).  These run once per object or per
tomograph (set-up), never per bootstrap sample; the per-sample arithmetic lives in csrc/.
"""

from functools import lru_cache

import numpy as np
import scipy.linalg as la

_PAULI_1 = np.array(
    [[[1, 0], [0, 1]], [[0, 1], [1, 0]], [[0, -1j], [1j, 0]], [[1, 0], [0, -1]]], dtype=np.complex128
)
_SIGMA_I, _SIGMA_X, _SIGMA_Y, _SIGMA_Z = _PAULI_1


@lru_cache(maxsize=None)
def _pauli_tensor(n_qubits):
    out = _PAULI_1
    for _ in range(n_qubits - 1):
        k, s = out.shape[0], out.shape[1]
        out = np.einsum("iab,jcd->ijacbd", out, _PAULI_1).reshape(4 * k, 2 * s, 2 * s)
    out = np.ascontiguousarray(out)
    out.setflags(write=False)
    return out


def generate_pauli(n_qubits):
    """All 4^n Pauli strings as a (4^n, 2^n, 2^n) array; index i1*4^(n-1)+...+in, sigma order I,X,Y,Z
    (quantpy/routines.py:14-19)."""
    return _pauli_tensor(int(n_qubits))


def generate_single_entries(dim):
    """The dim^2 matrix units E_ij in row-major (i, j) order (quantpy/routines.py:22-31)."""
    eye = np.eye(dim * dim).reshape(dim * dim, dim, dim)
    return [unit.copy() for unit in eye]


def kron(A, B):
    """Kronecker product of two quantum objects (quantpy/routines.py:34-36)."""
    return A.kron(B)


def join_gates(gates):
    """Compose gates applied left to right: join_gates([a, b]) == b @ a (quantpy/routines.py:39-44)."""
    total = gates[0]
    for gate in gates[1:]:
        total = gate @ total
    return total


def _vec2mat(vector):
    """Inverse of column stacking (quantpy/routines.py:53-56)."""
    vector = np.asarray(vector)
    side = int(round(np.sqrt(vector.shape[-1])))
    return np.swapaxes(vector.reshape(vector.shape[:-1] + (side, side)), -1, -2)


def _mat2vec(matrix):
    """Column stacking vec(M)[c*s + r] = M[r, c] (quantpy/routines.py:59-61)."""
    matrix = np.asarray(matrix)
    return np.swapaxes(matrix, -1, -2).reshape(matrix.shape[:-2] + (-1,))


def _density(psi):
    """|psi><psi| (quantpy/routines.py:64-66)."""
    psi = np.asarray(psi, dtype=np.complex128).reshape(-1)
    return np.outer(psi, psi.conj())


def _left_inv(A):
    """(A^T A)^-1 A^T with a plain (non-conjugating) transpose, as quantpy/routines.py:69-71.
    Computed once per tomograph and uploaded; the reference recomputes it on every call."""
    A = np.asarray(A)
    return la.inv(A.T @ A) @ A.T


def _out_ptrace_oper(n_qubits):
    """(d^2, d^4) matrix taking vec(Choi) to vec(Tr_out Choi) (quantpy/routines.py:47-50)."""
    d = 2**n_qubits
    op = np.zeros((d, d, d, d, d, d))  # [i, j | c_in, c_out, r_in, r_out] with vec index c*s + r
    for a in range(d):
        for i in range(d):
            for j in range(d):
                op[i, j, j, a, i, a] = 1.0
    # vec(rho)[j*d + i] = rho[i, j]
    return op.transpose(1, 0, 2, 3, 4, 5).reshape(d * d, d**4)
