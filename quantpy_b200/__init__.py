"""quantpy_b200 -- B200-native drop-in for the tomography-bootstrap hot path of nordmtr/quantpy.

    import quantpy_b200 as qp

exposes the names of quantpy/__init__.py that live on that path (Qobj, Channel, StateTomograph,
ProcessTomograph, BootstrapStateInterval, BootstrapProcessInterval, distances, POVM generator).
Importing the package needs neither a GPU nor the compiled library; calling anything that does
arithmetic on the path (experiment, point_estimate, the bootstrap intervals) needs both and raises
quantpy_b200.NativeError otherwise -- there is no CPU fallback.
"""

from . import basis, channel, operator, parallel  # noqa: F401
from ._native import NativeError  # noqa: F401
from .base_quantum import BaseQuantum  # noqa: F401
from .channel import Channel  # noqa: F401
from .geometry import hs_dst, if_dst, product, trace_dst  # noqa: F401
from .measurements import generate_measurement_matrix  # noqa: F401
from .operator import Operator  # noqa: F401
from .qobj import Qobj  # noqa: F401
from .routines import generate_pauli, join_gates, kron  # noqa: F401
from .tomography.interval import (  # noqa: F401
    BootstrapProcessInterval,
    BootstrapStateInterval,
    HolderInterval,
    MHMCProcessInterval,
    MHMCStateInterval,
    MomentFidelityProcessInterval,
    MomentFidelityStateInterval,
    MomentInterval,
    PolytopeProcessInterval,
    PolytopeStateInterval,
    SugiyamaInterval,
)
from .tomography.process import ProcessTomograph  # noqa: F401
from .tomography.state import StateTomograph  # noqa: F401

__version__ = "0.1.0"
