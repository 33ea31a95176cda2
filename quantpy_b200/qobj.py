"""Qobj: density matrices / observables with lazy matrix <-> Bloch conversion.

Host mirror of quantpy/qobj.py (same constructor, attributes and methods).  Qobj is the data
carrier of the hot path, not part of its arithmetic; the conversions are vectorised contractions
with the cached Pauli tensor (the reference loops over 4^n Pauli strings in Python, qobj.py:109-135).
"""

import math
import sys

import numpy as np
import scipy.linalg as la

from .base_quantum import BaseQuantum
from .routines import _density, generate_pauli


class Qobj(BaseQuantum):
    """Quantum state or Hermitian operator.

    data : 2-D array-like -> matrix; 1-D -> Bloch vector (a vector of length 4^n - 1 gets the
    identity coefficient 1/2^n prepended, qobj.py:94-99); 1-D with is_ket=True -> pure state.
    """

    def __init__(self, data, is_ket=False):
        if isinstance(data, Qobj):
            self.n_qubits = data.n_qubits
            self._matrix = None if data._matrix is None else np.array(data._matrix)
            self._bloch = None if data._bloch is None else np.array(data._bloch)
            return
        data = np.array(_density(data) if is_ket else data)
        self._matrix = self._bloch = None
        if data.ndim == 2:
            self._matrix = data
            self.n_qubits = int(np.log2(data.shape[0]))
        elif data.ndim == 1:
            exact = math.log2(data.shape[0]) / 2
            self.n_qubits = math.ceil(exact)
            if exact.is_integer():
                self._bloch = data
            else:
                dim = 2**self.n_qubits
                self._bloch = np.concatenate(([1.0 / dim], data))
        else:
            raise ValueError("Invalid data format")

    # -- representations ------------------------------------------------------------------------
    @property
    def matrix(self):
        if self._matrix is None:
            self._matrix = np.tensordot(self._bloch, generate_pauli(self.n_qubits), axes=1)
        return self._matrix

    @matrix.setter
    def matrix(self, data):
        self._matrix = np.array(data)
        self._bloch = None

    @property
    def bloch(self):
        if self._bloch is None:
            dim = 2**self.n_qubits
            paulis = generate_pauli(self.n_qubits)
            # Re Tr(sigma_i M^dagger) / 2^n   (qobj.py:131-134 with geometry.product)
            self._bloch = np.real(np.einsum("iab,ab->i", paulis, np.conj(self._matrix))) / dim
        return self._bloch

    @bloch.setter
    def bloch(self, data):
        self._bloch = np.array(data)
        self._matrix = None

    @property
    def _types(self):
        """Names of the cached representations (kept for code that introspects like the reference)."""
        return {name for name, val in (("matrix", self._matrix), ("bloch", self._bloch)) if val is not None}

    # -- operations -----------------------------------------------------------------------------
    def ptrace(self, keep=(0,)):
        """Partial trace keeping the listed qubits (qobj.py:145-165)."""
        keep = [int(k) for k in np.atleast_1d(keep)]
        n = self.n_qubits
        tensor = self.matrix.reshape([2] * (2 * n))
        rows = list(range(n))
        cols = [n + q if q in keep else q for q in range(n)]
        traced = np.einsum(tensor, rows + cols)
        side = 2 ** len(keep)
        return Qobj(traced.reshape(side, side))

    def schmidt(self):
        """SVD of the ket reshaped as a bipartite matrix (qobj.py:167-184)."""
        side = 2 ** int(self.n_qubits / 2)
        return la.svd(np.reshape(self.ket(), (side, side)))

    def eig(self):
        return la.eig(self.matrix)

    def is_density_matrix(self, verbose=True):
        m = self.matrix
        checks = (
            ("Non-hermitian", np.allclose(m, m.T.conj())),
            ("Non-positive", np.allclose(np.minimum(np.real(self.eig()[0]), 0), 0)),
            ("Trace is not 1", np.allclose(np.trace(m), 1)),
        )
        if verbose:
            for message, ok in checks:
                if not ok:
                    print(message, file=sys.stderr)
        return all(ok for _, ok in checks)

    def trace(self):
        return np.trace(self.matrix)

    def impurity(self):
        return 1 - (self @ self).trace()

    def is_pure(self):
        return bool(np.allclose(self.impurity(), 0)) and self.is_density_matrix()

    def ket(self):
        if not self.is_pure():
            raise ValueError("Quantum object is not pure")
        return self.eig()[1][:, 0]

    def __repr__(self):
        return "Quantum object\n" + repr(self.matrix)


def fully_mixed(n_qubits=1):
    dim = 2**n_qubits
    return Qobj(np.eye(dim, dtype=np.complex128) / dim)


# noinspection PyPep8Naming
def GHZ(n_qubits=3):
    ket = np.zeros(2**n_qubits)
    ket[0] = ket[-1] = 1 / np.sqrt(2)
    return Qobj(ket, is_ket=True)


def zero(n_qubits=1):
    ket = np.zeros(2**n_qubits)
    ket[0] = 1
    return Qobj(ket, is_ket=True)
