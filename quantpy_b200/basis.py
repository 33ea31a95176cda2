"""Basis of a Euclidean space of quantum objects (host mirror of quantpy/basis.py).  Used once per
ProcessTomograph to express matrix units through the input states; not on the per-sample path."""

import numpy as np
import scipy.linalg as la

from .geometry import product


class Basis:
    def __init__(self, elements, inner_product="trace"):
        self.elements = elements
        self.dim = len(elements)
        self.inner_product = product if inner_product == "trace" else inner_product
        self.gram = np.array(
            [[self.inner_product(a, b) for b in elements] for a in elements], dtype=np.complex128
        ).reshape(self.dim, self.dim)

    def decompose(self, obj):
        """Coefficients c with obj = sum_i c_i e_i (quantpy/basis.py:31-34)."""
        rhs = np.array([self.inner_product(e, obj) for e in self.elements], dtype=np.complex128)
        return np.conj(la.solve(self.gram, rhs))

    def compose(self, vector):
        """sum_i vector_i e_i (quantpy/basis.py:36-38)."""
        total = self.elements[0] * complex(vector[0])
        for element, coef in zip(self.elements[1:], vector[1:]):
            total = total + element * complex(coef)
        return total

    def __repr__(self):
        return "Basis object\n" + repr(self.elements)
