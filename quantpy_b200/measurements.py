"""POVM tables in the Pauli (Bloch) basis -- host mirror of quantpy/measurements.py.

One-time set-up on the host; the result is uploaded to the device by the tomograph plan.
"""

import numpy as np

_AXES = {
    "x+": (1, 1, 0, 0), "x-": (1, -1, 0, 0),
    "y+": (1, 0, 1, 0), "y-": (1, 0, -1, 0),
    "z+": (1, 0, 0, 1), "z-": (1, 0, 0, -1),
}


def _rows(*names):
    return np.array([_AXES[name] for name in names])


def _proto_povm(name):
    """Single-qubit proto-POVM as a (P, O, 4) array (quantpy/measurements.py:36-72)."""
    if name == "proj":
        return _rows("x+", "x-", "y+", "y-", "z+", "z-")[None] / 6
    if name == "proj-set":
        return np.stack([_rows("x+", "x-"), _rows("y+", "y-"), _rows("z+", "z-")]) / 2
    if name == "proj4":
        return _rows("x+", "y+", "z+", "z-")[None] / 4
    if name == "sic":
        s = 1 / np.sqrt(3)
        return np.array([[1, s, s, s], [1, s, -s, -s], [1, -s, s, -s], [1, -s, -s, s]])[None] / 4
    raise ValueError("Incorrect string shortcut for argument `povm`")


def generate_measurement_matrix(povm="proj", n_qubits=1):
    """Return the (P, O, 4^n) POVM tensor whose rows are Pauli coefficients of the POVM elements.

    povm : 'proj' | 'proj-set' | 'proj4' | 'sic', or an array: (*, 4) / (*, *, 4) single-qubit tables
    are tensored n_qubits times, (*, 4^n) / (*, *, 4^n) tables are returned as given (2-D tables gain a
    leading axis).  Same contract and error behaviour as quantpy/measurements.py:4-94.
    """
    if isinstance(povm, str):
        proto = _proto_povm(povm)
    elif isinstance(povm, np.ndarray):
        width = povm.shape[-1]
        if width == 4:
            proto = povm if povm.ndim == 3 else povm[None]
        elif width == 4**n_qubits:
            return povm if povm.ndim == 3 else povm[None]
        else:
            raise ValueError("Incorrect POVM matrix")
    else:
        raise ValueError("Incorrect value for argument `povm`")
    full = proto
    for _ in range(n_qubits - 1):
        full = np.kron(full, proto)
    return full
