"""Arithmetic shared by Qobj / Operator (host mirror of quantpy/base_quantum.py:8-89).

Every object exposes a `.matrix`; the operators below are installed from small factories instead of being spelled
out one by one: matrix (x) matrix -> new object, object (x) scalar -> new object, and the in-place forms.
Error behaviour follows the reference: scaling by anything but a scalar raises ValueError."""

import copy as _copy
import operator as _op
from abc import ABC, abstractmethod

import numpy as np

_SCALARS = (int, float, complex, np.integer, np.floating, np.complexfloating)


def _check_scalar(value, what, verdict):
    if not isinstance(value, _SCALARS):
        raise ValueError(f"Only {what} by a scalar is {verdict}")


class BaseQuantum(ABC):
    """Matrix-backed quantum object with elementwise arithmetic."""

    __hash__ = None
    __array_ufunc__ = None  # numpy scalars defer to __rmul__ instead of broadcasting over us

    @abstractmethod
    def __repr__(self):
        ...

    def _wrap(self, matrix):
        return type(self)(matrix)

    T = property(lambda self: self._wrap(self.matrix.T), doc="transpose")
    H = property(lambda self: self._wrap(self.matrix.conj().T), doc="adjoint")

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

    def __neg__(self):
        return self._wrap(-self.matrix)


def _with_object(fn):
    def method(self, other):
        return self._wrap(fn(self.matrix, other.matrix))

    return method


def _with_scalar(fn, what):
    def method(self, other):
        _check_scalar(other, what, "allowed")
        return self._wrap(fn(self.matrix, other))

    return method


def _in_place(fn, what=None):
    def method(self, other):
        if what is None:
            self.matrix = fn(self.matrix, other.matrix)
        else:
            _check_scalar(other, what, "supported")
            self.matrix = fn(self.matrix, other)
        return self

    return method


for _name, _fn in (("__matmul__", _op.matmul), ("__add__", _op.add), ("__sub__", _op.sub)):
    setattr(BaseQuantum, _name, _with_object(_fn))
for _name, _fn in (("__iadd__", _op.add), ("__isub__", _op.sub)):
    setattr(BaseQuantum, _name, _in_place(_fn))
BaseQuantum.__mul__ = BaseQuantum.__rmul__ = _with_scalar(_op.mul, "multiplication")
BaseQuantum.__truediv__ = _with_scalar(_op.truediv, "division")
BaseQuantum.__imul__ = _in_place(_op.mul, "multiplication")
BaseQuantum.__itruediv__ = BaseQuantum.__idiv__ = _in_place(_op.truediv, "division")
