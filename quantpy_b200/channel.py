"""Channel: quantum channels as Choi matrix, Kraus list or Python callable.

Host mirror of quantpy/channel.py.  On the tomography path a Channel is an input type
(`ProcessTomograph(channel, ...)`) and the return type of `point_estimate`; applying it to the
S input states happens once per tomograph, not per bootstrap sample.
"""

import sys

import numpy as np

from .base_quantum import BaseQuantum
from .operator import H, Operator, Z, _choi_to_kraus
from .qobj import Qobj, fully_mixed
from .routines import generate_single_entries

_SCALARS = (int, float, complex, np.integer, np.floating, np.complexfloating)


class Channel(BaseQuantum):
    """data : callable (needs n_qubits) | 2-D array or Qobj (Choi matrix) | list of Operators (Kraus)."""

    def __init__(self, data, n_qubits=None):
        self._choi = self._kraus = self._func = None
        if isinstance(data, Channel):
            self.n_qubits = data.n_qubits
            self._choi = None if data._choi is None else Qobj(data._choi)
            self._kraus = None if data._kraus is None else list(data._kraus)
            self._func = data._func
        elif callable(data):
            if n_qubits is None:
                raise ValueError("`n_qubits` argument is compulsory when using init with function")
            self._func = data
            self.n_qubits = n_qubits
        elif isinstance(data, (np.ndarray, Qobj)):
            self._choi = Qobj(data)
            self.n_qubits = int(self._choi.n_qubits / 2)
        elif isinstance(data, list):
            self._kraus = data
            self.n_qubits = data[0].n_qubits
        else:
            raise ValueError("Invalid data format")

    def set_func(self, data, n_qubits):
        self._choi = self._kraus = None
        self._func = data
        self.n_qubits = n_qubits

    @property
    def _types(self):
        pairs = (("choi", self._choi), ("kraus", self._kraus), ("func", self._func))
        return {name for name, val in pairs if val is not None}

    @property
    def choi(self):
        """sum_ij E_ij (x) Phi(E_ij) (quantpy/channel.py:95-103)."""
        if self._choi is None:
            side = 4**self.n_qubits
            total = np.zeros((side, side), dtype=np.complex128)
            for unit in generate_single_entries(2**self.n_qubits):
                total += np.kron(unit, self.transform(unit).matrix)
            self._choi = Qobj(total)
        return self._choi

    @choi.setter
    def choi(self, data):
        self._choi = data if isinstance(data, Qobj) else Qobj(data)
        self._kraus = self._func = None
        self.n_qubits = int(self._choi.n_qubits / 2)

    @property
    def matrix(self):
        """Arithmetic on channels acts on the Choi matrix."""
        return self.choi.matrix

    @matrix.setter
    def matrix(self, data):
        self.choi = data

    @property
    def kraus(self):
        if self._kraus is None:
            self._kraus = _choi_to_kraus(self.choi)
        return self._kraus

    @kraus.setter
    def kraus(self, data):
        if not isinstance(data, list):
            raise ValueError("Invalid data format")
        self._kraus = data
        self._choi = self._func = None
        self.n_qubits = data[0].n_qubits

    def transform(self, state):
        """Apply the channel: Kraus sum, the callable, or Tr_in[(rho^T (x) I) Choi] (channel.py:131-142)."""
        if not isinstance(state, Qobj):
            state = Qobj(state)
        if self._kraus is not None:
            out = self._kraus[0].transform(state)
            for oper in self._kraus[1:]:
                out = out + oper.transform(state)
            return out
        if self._func is not None:
            return self._func(state)
        d = 2**self.n_qubits
        choi4 = self._choi.matrix.reshape(d, d, d, d)
        return Qobj(np.einsum("ij,iajb->ab", state.matrix, choi4))

    def is_cptp(self, atol=1e-5, verbose=True):
        d = 2**self.n_qubits
        rho_in = np.einsum("iaja->ij", self.choi.matrix.reshape(d, d, d, d))
        tp = np.allclose(rho_in, np.eye(d), atol=atol)
        cp = np.allclose(np.minimum(np.real(self.choi.eig()[0]), 0), 0, atol=atol)
        if verbose and not tp:
            print("Not trace-preserving", file=sys.stderr)
        if verbose and not cp:
            print("Not completely positive", file=sys.stderr)
        return bool(tp and cp)

    def _wrap(self, matrix):
        return Channel(matrix)

    def __repr__(self):
        return "Quantum channel with Choi matrix\n" + repr(self.choi.matrix)


def depolarizing(p=1, n_qubits=1):
    """rho -> p Tr(rho) I/d + (1-p) rho (quantpy/channel.py:232-236)."""
    return Channel(lambda rho: p * rho.trace() * fully_mixed(n_qubits) + (1 - p) * rho, n_qubits)


def dephasing(p=1, n_qubits=1):
    """rho -> (1-p) rho + p Z rho Z."""
    return Channel(lambda rho: p * Z.transform(rho) + (1 - p) * rho, n_qubits)


def amplitude_damping(gamma):
    return Channel([
        np.sqrt(gamma) * Operator([[0, 1], [0, 0]]),
        Operator([[1, 0], [0, 0]]) + np.sqrt(1 - gamma) * Operator([[0, 0], [0, 1]]),
    ])


def walsh_hadamard(n_qubits):
    gate = H
    for _ in range(n_qubits - 1):
        gate = gate.kron(H)
    return gate.as_channel()


def depolarize(channel, p):
    return Channel(
        lambda rho: (1 - p) * channel.transform(rho) + p * rho.trace() * fully_mixed(channel.n_qubits),
        channel.n_qubits,
    )
