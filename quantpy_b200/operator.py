"""Operator: unitary gates acting by conjugation (host mirror of the part of quantpy/operator.py
that Channel needs; the gate zoo is API sugar outside the tomography hot path)."""

import numpy as np

from .base_quantum import BaseQuantum
from .qobj import Qobj
from .routines import _SIGMA_I, _SIGMA_X, _SIGMA_Y, _SIGMA_Z, _vec2mat


class Operator(BaseQuantum):
    def __init__(self, data):
        if isinstance(data, Operator):
            data = data.matrix
        self.matrix = data

    @property
    def matrix(self):
        return self._matrix

    @matrix.setter
    def matrix(self, data):
        self._matrix = np.array(data, dtype=np.complex128)
        self.n_qubits = int(np.log2(self._matrix.shape[0]))

    def transform(self, state):
        """U rho U^dagger (quantpy/operator.py:61-63)."""
        rho = state.matrix if isinstance(state, Qobj) else np.asarray(state)
        return Qobj(self._matrix @ rho @ self._matrix.conj().T)

    def as_channel(self):
        from .channel import Channel

        return Channel(self.transform, self.n_qubits)

    def trace(self):
        return np.trace(self._matrix)

    def __repr__(self):
        return "Quantum Operator\n" + repr(self._matrix)


def _choi_to_kraus(choi):
    """Kraus operators from the eigen-decomposition of a Choi matrix."""
    vals, vecs = np.linalg.eigh(choi.matrix)
    ops = []
    for val, vec in zip(vals, vecs.T):
        if val > 1e-12:
            ops.append(Operator(_vec2mat(vec * np.sqrt(val)).T))
    return ops


Id = Operator(_SIGMA_I)
X = Operator(_SIGMA_X)
Y = Operator(_SIGMA_Y)
Z = Operator(_SIGMA_Z)
H = Operator(np.array([[1, 1], [1, -1]]) / np.sqrt(2))
S = Operator(np.diag([1, 1j]))
T = Operator(np.diag([1, np.exp(1j * np.pi / 4)]))
CNOT = Operator(np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]]))
CZ = Operator(np.diag([1, 1, 1, -1]))
SWAP = Operator(np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]))
