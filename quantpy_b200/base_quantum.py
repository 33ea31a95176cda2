"""Arithmetic shared by Qobj / Operator (host mirror of quantpy/base_quantum.py)."""

import copy as _copy
from abc import ABC, abstractmethod

import numpy as np

_SCALARS = (int, float, complex, np.integer, np.floating, np.complexfloating)


class BaseQuantum(ABC):
    """Matrix-backed quantum object with elementwise arithmetic.  Subclasses expose `.matrix`."""

    @abstractmethod
    def __repr__(self):
        ...

    def _wrap(self, matrix):
        return type(self)(matrix)

    @property
    def T(self):
        return self._wrap(self.matrix.T)

    @property
    def H(self):
        return self._wrap(self.matrix.conj().T)

    def conj(self):
        return self._wrap(self.matrix.conj())

    def copy(self):
        return _copy.deepcopy(self)

    def kron(self, other):
        return self._wrap(np.kron(self.matrix, other.matrix))

    def __eq__(self, other):
        return np.array_equal(self.matrix, other.matrix)

    def __ne__(self, other):
        return not self == other

    __hash__ = None
    __array_ufunc__ = None  # numpy scalars defer to __rmul__ instead of broadcasting over us

    def __neg__(self):
        return self._wrap(-self.matrix)

    def __matmul__(self, other):
        return self._wrap(self.matrix @ other.matrix)

    def __add__(self, other):
        return self._wrap(self.matrix + other.matrix)

    def __sub__(self, other):
        return self._wrap(self.matrix - other.matrix)

    @staticmethod
    def _need_scalar(value, what):
        if not isinstance(value, _SCALARS):
            raise ValueError(f"Only {what} by a scalar is allowed")

    def __mul__(self, other):
        self._need_scalar(other, "multiplication")
        return self._wrap(self.matrix * other)

    __rmul__ = __mul__

    def __truediv__(self, other):
        self._need_scalar(other, "division")
        return self._wrap(self.matrix / other)

    def __iadd__(self, other):
        self.matrix = self.matrix + other.matrix
        return self

    def __isub__(self, other):
        self.matrix = self.matrix - other.matrix
        return self

    def __imul__(self, other):
        self._need_scalar(other, "multiplication")
        self.matrix = self.matrix * other
        return self

    def __itruediv__(self, other):
        self._need_scalar(other, "division")
        self.matrix = self.matrix / other
        return self

    __idiv__ = __itruediv__
