"""StateTomograph -- drop-in for quantpy/tomography/state.py on the B200 path.

Same constructor, attributes (`state`, `dst`, `povm_matrix`, `results`, `flat_results`,
`n_measurements`, `reconstructed_state`) and methods (`experiment`, `point_estimate`) as the
reference class.  The per-sample arithmetic runs in CUDA kernels through quantpy_b200.engine:

  experiment      -> qpb_povm_probs + qpb_multinomial      (state.py:99-128)
  point_estimate  -> qpb_lin_project / qpb_mle_rrr         (state.py:143-229, 267-273)

`method='mle'` is the iterative R.rho.R maximum-likelihood update asked for by the task statement,
not the reference's SciPy BFGS over a Cholesky factor: `max_iter` caps the number of R.rho.R
iterations and `tol` is the Frobenius norm of the last step (see DESIGN.md "MLE semantics").
`'mle-constr'` (the reference's SLSQP variant of the same likelihood under Tr rho = 1, state.py:231-254) has the
same maximiser over physical states; it is accepted and runs the same R.rho.R kernel.

Batched extensions used by the bootstrap: `sample_counts`, `point_estimate_batch`.
"""

import numpy as np

from .. import _native as nt
from .. import engine
from ..geometry import hs_dst, if_dst, trace_dst
from ..measurements import generate_measurement_matrix
from ..qobj import Qobj

_DST_BY_NAME = {"hs": hs_dst, "trace": trace_dst, "if": if_dst}


def resolve_dst(dst):
    if isinstance(dst, str):
        if dst not in _DST_BY_NAME:
            raise ValueError("Invalid value for argument `dst`")
        return _DST_BY_NAME[dst]
    return dst


def dst_kind(dst_fn):
    """Name of the built-in distance behind a dst callable, or None for a user-supplied measure."""
    for name, fn in _DST_BY_NAME.items():
        if dst_fn is fn:
            return name
    return None


class StateTomograph:
    def __init__(self, state, dst="hs"):
        self.state = state
        self.dst = resolve_dst(dst)
        self._results = None

    # -- experiment ------------------------------------------------------------------------------
    def _shot_vector(self, n_measurements, number_of_povms):
        if np.issubdtype(type(n_measurements), np.integer):
            return np.ones(number_of_povms) * n_measurements
        if len(n_measurements) != number_of_povms:
            raise ValueError("Wrong length for argument `n_measurements`")
        return n_measurements

    def sample_counts(self, n_samples, n_measurements, povm="proj-set", seed=None, offset=0, device=False):
        """Simulate `n_samples` independent experiments at once -> counts [n_samples, P, O].

        Sample b is drawn from the Philox stream keyed by (seed, offset + b); `seed=None` takes the key
        from NumPy's global RNG.  device=True returns the int32 CUDA tensor without copying it back."""
        povm_matrix = generate_measurement_matrix(povm, self.state.n_qubits)
        shots = self._shot_vector(n_measurements, povm_matrix.shape[0])
        plan = engine.state_plan(povm_matrix, shots)
        probs = plan.probabilities(self.state.bloch)[0]
        counts = plan.sample(probs, int(n_samples), engine.next_seed() if seed is None else int(seed), int(offset))
        return counts if device else counts.cpu().numpy().astype(np.int64)

    def experiment(self, n_measurements, povm="proj-set", warm_start=False):
        """Simulate one tomography run: multinomial shot counts for every POVM (state.py:71-128)."""
        povm_matrix = generate_measurement_matrix(povm, self.state.n_qubits)
        n_measurements = self._shot_vector(n_measurements, povm_matrix.shape[0])
        results = self.sample_counts(1, n_measurements, povm_matrix)[0]
        if warm_start:
            old_total, new_total = np.sum(self.n_measurements), np.sum(n_measurements)
            self.povm_matrix = np.vstack((self.povm_matrix * old_total, povm_matrix * new_total)) / (
                old_total + new_total
            )
            merged = np.hstack((self.n_measurements, n_measurements))
            self.results = np.vstack((self.results, results))
            self.n_measurements = merged
        else:
            self.povm_matrix = povm_matrix
            self.results = results
            self.n_measurements = np.asarray(n_measurements)

    @property
    def flat_results(self):
        return self.results.flatten()

    @property
    def results(self):
        return self._results

    @results.setter
    def results(self, results):
        self._results = results
        self.n_measurements = results.sum(-1)

    # -- reconstruction --------------------------------------------------------------------------
    def _plan(self):
        return engine.state_plan(self.povm_matrix, self.n_measurements)

    def point_estimate_batch(self, counts, method="lin", physical=True, init="lin", max_iter=100, tol=1e-3,
                             return_iters=False):
        """Reconstruct a batch of count tables [B, P, O] measured with this tomograph's POVM and shots.
        Returns complex matrices [B, d, d] (and the R.rho.R iteration counts when asked)."""
        torch = nt.torch_cuda()
        if method == "mle-constr":  # same likelihood, same feasible set: R.rho.R keeps Tr rho = 1 and rho >= 0
            method = "mle"
        if method not in ("lin", "mle"):
            raise ValueError("Invalid value for argument `method`")
        plan = self._plan()
        if not torch.is_tensor(counts):
            counts = nt.to_device(np.asarray(counts).reshape(-1, plan.K), torch.int32)
        rho, iters = plan.estimate(counts, method, physical, init, max_iter, tol)
        out = nt.complex_to_host(rho)
        if return_iters:
            return out, (None if iters is None else iters.cpu().numpy())
        return out

    def point_estimate(self, method="lin", physical=True, init="lin", max_iter=100, tol=1e-3):
        """Reconstruct the density matrix from `results` (state.py:143-189).

        method : 'lin' (linear inversion) | 'mle' | 'mle-constr' (both: iterative maximum likelihood)
        physical : project the 'lin' estimate onto physical states (clip eigenvalues at 1e-15, renormalise)
        init : start of the MLE iteration, 'lin' (physical linear-inversion estimate) | 'mixed'
        max_iter, tol : iteration cap and step-norm stopping threshold of the MLE iteration
        """
        matrix = self.point_estimate_batch(self.results[None], method, physical, init, max_iter, tol)[0]
        self.reconstructed_state = Qobj(matrix)
        return self.reconstructed_state
