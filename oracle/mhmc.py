"""Oracle (TEST INFRASTRUCTURE ONLY): Metropolis-Hastings likelihood sampling of quantpy/mhmc.py and
MHMCStateInterval (quantpy/tomography/interval.py:689-759), restated in NumPy.

One chain, sequential, exactly as the reference runs it: the state is the packed Cholesky vector x of
routines.py:84-101 (|x| = 1, so Tr L L^dagger = 1), the proposal is x' = (x + step*delta)/|x + step*delta|
(`normalized_update`, mhmc.py:117-119) with delta ~ N(0, I), the target is exp(-nll(x)) with the reference's
frequency-weighted log-likelihood (state.py:217-229), and the jump is treated as symmetric.  Random numbers are
drawn from the legacy global NumPy stream in the reference's order (burn-in: all deltas, then all uniforms;
sampling: all deltas, then all uniforms) unless the caller passes them in.

Parity: pinned by tests/golden/mhmc.npz -- replaying np.random.seed reproduces the reference's sorted distances.
"""

import numpy as np
from scipy.stats import multivariate_normal

from . import distances as odist
from . import state as ostate


def log_target(x, counts, povm, n_meas):
    """-StateTomograph._nll(x): sum_k f_k log(p_k + 1e-10), rho = L L^dagger / Tr.  state.py:217-229."""
    m = ostate._chol_unpack(np.asarray(x, dtype=float))
    return -ostate.neg_log_likelihood(m / np.trace(m), counts, povm, n_meas)


def draw(dim, size):
    """The reference's draws for `size` steps: jump_distr.rvs then np.random.rand.  mhmc.py:70-71, 89-90."""
    deltas = multivariate_normal(mean=np.zeros(dim)).rvs(size=size)
    return np.reshape(deltas, (size, dim)), np.random.rand(size)


def chain(x_init, logpdf, n_samples, step=0.01, burn_steps=100, thinning=1, burn_draws=None, draws=None):
    """MHMC(...).sample(n_samples, thinning) on a fresh chain.  mhmc.py:48-112.
    Returns (samples [n_samples, dim], acceptance_rate, x_final)."""
    x = np.asarray(x_init, dtype=float).copy()
    dim = len(x)
    cur = logpdf(x)

    def advance(x, cur, delta, u):
        prop = x + step * delta
        prop = prop / np.linalg.norm(prop)
        new = logpdf(prop)
        if u <= np.exp(new - cur):
            return prop, new, True
        return x, cur, False

    deltas, us = draw(dim, burn_steps) if burn_draws is None else burn_draws
    for i in range(burn_steps):
        x, cur, _ = advance(x, cur, deltas[i], us[i])
    total = n_samples * thinning
    deltas, us = draw(dim, total) if draws is None else draws
    samples = np.zeros((n_samples, dim))
    accepted = 0
    for i in range(total):
        x, cur, ok = advance(x, cur, deltas[i], us[i])
        accepted += ok
        if i % thinning == 0:
            samples[i // thinning] = x
    return samples, accepted / total, x


def state_interval(counts, povm, n_meas, state_matrix, dst="hs", n_points=1000, step=0.01, burn_steps=1000,
                   thinning=1, burn_draws=None, draws=None, return_chain=False):
    """MHMCStateInterval.setup: sorted distances between the chain's L L^dagger and the centre state.
    interval.py:737-759."""
    x0 = ostate._chol_pack(np.asarray(state_matrix, dtype=np.complex128))
    samples, rate, x_final = chain(x0, lambda x: log_target(x, counts, povm, n_meas), n_points, step, burn_steps,
                                   thinning, burn_draws, draws)
    fn = odist.BY_NAME[dst]
    dist = np.sort([fn(ostate._chol_unpack(s), state_matrix) for s in samples])
    if return_chain:
        return dist, samples, rate, x_final
    return dist
