"""Polytope-coverage oracle (test infrastructure only).

Restates quantpy/tomography/polytopes/utils.py:4-27 (count_confidence, count_delta) and the per-trial
body of quantpy/tomography/polytopes/verification.py:9-78 (test_qst, test_qpt) on explicit count tables,
so that the CUDA kernel can be compared trial by trial instead of only through Monte-Carlo averages.
"""

import numpy as np

EPS = 1e-15


def count_confidence(delta, frequencies, n_measurements):
    """utils.py:4-13.  frequencies: (..., P, O) clipped to [EPS, 1-EPS]; n_measurements: (P,)."""
    f = np.asarray(frequencies, dtype=float)
    shifted = np.clip(f + delta, EPS, 1 - EPS)
    with np.errstate(divide="ignore", invalid="ignore"):
        kl = f * np.log(f / shifted) + (1 - f) * np.log((1 - f) / (1 - shifted))
    kl = np.where(shifted < 1 - EPS, kl, np.inf)
    eps = np.exp(-np.asarray(n_measurements, dtype=float)[:, None] * kl)
    eps = np.where(np.abs(f - 1) < 2 * EPS, 0, eps)
    return np.prod(np.maximum(1 - np.sum(eps, axis=-1), 0))


def count_delta(target_cl, frequencies, n_measurements):
    """utils.py:16-27: bisection on [1e-10, 1] down to a width of 1e-10; returns the last midpoint."""
    left, right, delta = 1e-10, 1.0, None
    while right - left > 1e-10:
        delta = (left + right) / 2
        if count_confidence(delta, frequencies, n_measurements) < target_cl + 1e-10:
            left = delta
        else:
            right = delta
    return delta


def clipped_frequencies(counts, n_measurements):
    """verification.py:29 / :65-67."""
    counts = np.asarray(counts, dtype=float)
    n = np.asarray(n_measurements, dtype=float)
    return np.clip(counts / n[:, None], EPS, 1 - EPS)


def true_probabilities_state(povm, bloch):
    """p_k of the true state in the form verification.py:17-25 builds it: povm0 + A.bloch[1:]
    (for equal shots the shot weighting times P cancels)."""
    povm = np.asarray(povm, dtype=float)
    dim = int(round(np.sqrt(povm.shape[-1])))
    flat = povm.reshape(-1, povm.shape[-1])
    return flat[:, 0] + (flat[:, 1:] * dim) @ np.asarray(bloch)[1:]


def trial_state(counts, n_measurements, p_true, conf_levels, clip_b=True):
    """One trial of test_qst (verification.py:27-35) or, with clip_b=False and the (S,P,O) tables flattened
    to (S*P, O), of test_qpt (:64-77).  Returns (deltas, contained) per confidence level."""
    f = clipped_frequencies(counts, n_measurements)
    deltas, inside = [], []
    for cl in conf_levels:
        d = count_delta(cl, f, n_measurements)
        b = np.hstack(f) + d
        if clip_b:
            b = np.clip(b, EPS, 1 - EPS)
        deltas.append(d)
        inside.append(bool(np.min(b - p_true) > -EPS))
    return np.array(deltas), np.array(inside)
