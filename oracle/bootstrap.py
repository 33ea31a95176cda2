"""Bootstrap-loop oracle (test infrastructure only).

Serial restatement of quantpy/tomography/interval.py:583-612 (state) and :658-685
(process): resample counts from the centre object, reconstruct, measure the distance
to the centre, sort, and interpolate the quantile function.  This is also what
``bench.py --impl reference`` times as the reference's CPU path (the reference is
pure Python and cannot be shipped to the GPU box).
"""

import numpy as np

from . import distances as odist
from . import process as oproc
from . import state as ostate
from .pauli import matrix_to_bloch


def quantile_function(sorted_dist):
    """interval.py:610-612: interp1d(linspace(0,1,N), sorted distances) (linear)."""
    grid = np.linspace(0, 1, len(sorted_dist))

    def cl_to_dist(levels):
        levels = np.asarray(levels, dtype=float)
        if np.any(levels < 0) or np.any(levels > 1):
            raise ValueError("A value in x_new is outside the interpolation range.")
        return np.interp(levels, grid, sorted_dist)

    return cl_to_dist


def state_point_estimate(counts, povm, n_meas, method="lin", physical=True, init="lin",
                         max_iter=100, tol=1e-3, mle="bfgs"):
    """Dispatcher of state.py:143-189 over the oracle estimators."""
    if method == "lin":
        return ostate.lin_estimate(counts, povm, n_meas, physical=physical)
    if method == "mle":
        if mle == "bfgs":
            return ostate.mle_bfgs(counts, povm, n_meas, init=init, max_iter=max_iter, tol=tol)
        return ostate.mle_rrr(counts, povm, n_meas, init=init, max_iter=max_iter, tol=tol)
    raise ValueError("Invalid value for argument `method`")


def bootstrap_state(centre, povm, n_meas, n_points, method="lin", physical=True, init="lin",
                    max_iter=100, tol=1e-3, dst="hs", mle="bfgs", rng=None, sort=True):
    """interval.py:598-610.  ``centre`` is the (d, d) state resampled from."""
    dist_fn = odist.BY_NAME[dst] if isinstance(dst, str) else dst
    bloch = matrix_to_bloch(centre)
    dist = np.empty(n_points)
    for i in range(n_points):
        counts = ostate.experiment(povm, bloch, n_meas, rng=rng)
        rho = state_point_estimate(counts, povm, n_meas, method, physical, init, max_iter, tol, mle)
        dist[i] = dist_fn(rho, centre)
    if sort:
        dist.sort()
    return dist


def bootstrap_process(centre_choi, inputs, povm, n_meas, n_points, method="lifp", cptp=True,
                      dst="hs", states_est_method="lin", states_physical=True, states_init="lin",
                      mle="bfgs", rng=None, sort=True):
    """interval.py:672-683."""
    dist_fn = odist.BY_NAME[dst] if isinstance(dst, str) else dst
    dist = np.empty(n_points)
    for i in range(n_points):
        counts = oproc.experiment(centre_choi, inputs, povm, n_meas, rng=rng)
        if method == "lifp":
            est = oproc.lifp_estimate(counts, inputs, povm, n_meas, cptp=cptp)
        elif method == "states":
            est = oproc.states_estimate(counts, inputs, povm, n_meas, cptp=cptp, method=states_est_method,
                                        physical=states_physical, init=states_init, mle=mle)
        else:
            raise ValueError("Incorrect value for argument `method`")
        dist[i] = dist_fn(est, centre_choi)
    if sort:
        dist.sort()
    return dist
