"""CPU oracle for the tomography-bootstrap hot path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy restatement of the algorithms on the hot path of
nordmtr/quantpy (state/process tomography bootstrap).  It exists so that the
CUDA kernels in ``quantpy_b200`` can be checked against an independent CPU
implementation.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The product package ``quantpy_b200`` never imports it and has no CPU fallback.

Parity status:

* ``lin``/projection/distances/lifp/CPTP restatements are PINNED against the
  reference itself: ``tools/make_golden.py`` imports ``/root/reference`` in the
  build container (with a stub for the absent ``cvxopt``) and stores its outputs
  in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every oracle
  function against them.  The reference ships no tests or golden vectors of its
  own, so the goldens are reference *outputs*, not reference
  fixtures.
* ``mle_bfgs`` restates the reference's BFGS/Cholesky estimator
  (quantpy/tomography/state.py:204-229) on top of SciPy and is pinned the same way.
* ``mle_rrr`` (the iterative R·rho·R update BASELINE.json's north_star asks for)
  has NO counterpart in the reference.  It is the specification of the GPU
  kernel; it is pinned to the reference only loosely (its likelihood is never
  worse than the reference's BFGS optimum, golden-checked).

Each function cites the reference file:line it follows.
"""

from .pauli import pauli_basis, bloch_to_matrix, matrix_to_bloch  # noqa: F401
