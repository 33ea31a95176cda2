"""Moment-interval oracle (test infrastructure only).

Restates quantpy/stats.py:5-53 (moments of the weighted squared l2 norm of f - p for multinomial
frequencies) and quantpy/tomography/interval.py:59-110 (MomentInterval).  The twelve six-operand
contractions of stats.py:32-47 are regrouped by POVM pair (a, b): with the O x O blocks W_ab = W[a,:,b,:],

    G_ab = f_a^T W_ab f_b            u_ab[j] = (f_a^T W_ab)_j            v_ab[i] = (W_ab f_b)_i
    T0 = sum_a diag(W_aa) . f_a      T1 = sum_a G_aa

    n   E|.|^2     = T0 - T1
    n^2 E|.|^4     = (T1 - T0)^2 + sum_ab [ G_ab G_ba + G_ab^2
                       - sum_j f_bj u_ab[j] v_ba[j] - sum_i f_ai v_ab[i] u_ba[i]
                       - sum_j f_bj u_ab[j]^2 - sum_i f_ai v_ab[i]^2
                       + sum_ij W_ab[i,j] (W_ba[j,i] + W_ab[i,j]) f_ai f_bj ]
"""

import numpy as np
import scipy.stats as sts

from . import state as ostate
from .pauli import n_qubits_from_D


def identity_weights(P, O):
    """stats.py:50-53."""
    w = np.zeros((P, O, P, O))
    for a in range(P):
        for i in range(O):
            w[a, i, a, i] = 1.0
    return w


def l2_moments(freq, n_trials, weights=None):
    """(mean, variance) of the weighted squared distance, stats.py:5-47."""
    f = np.asarray(freq, dtype=float)
    P, O = f.shape
    W = identity_weights(P, O) if weights is None else np.asarray(weights, dtype=float)
    G = np.einsum("aibj,ai,bj->ab", W, f, f)
    U = np.einsum("aibj,ai->abj", W, f)   # u_ab[j]
    V = np.einsum("aibj,bj->abi", W, f)   # v_ab[i]
    T0 = np.einsum("aiai,ai->", W, f)
    T1 = np.trace(G)
    rest = (np.sum(G * G.T) + np.sum(G * G)
            - np.einsum("bj,abj,baj->", f, U, V) - np.einsum("ai,abi,bai->", f, V, U)
            - np.einsum("bj,abj->", f, U**2) - np.einsum("ai,abi->", f, V**2)
            + np.einsum("aibj,bjai,ai,bj->", W, W, f, f) + np.einsum("aibj,ai,bj->", W**2, f, f))
    mean = (T0 - T1) / n_trials
    second = ((T1 - T0) ** 2 + rest) / n_trials**2
    return mean, second - mean**2


def state_weights(povm, counts):
    """interval.py:70-76, 89: weights of the squared Bloch-vector error."""
    povm = np.asarray(povm, dtype=float)
    P, O, D = povm.shape
    dim = 2 ** n_qubits_from_D(D)
    inv = ostate.left_inverse(povm.reshape(P * O, D)) / dim
    inv = inv.reshape(D, P, O)
    return np.einsum("aij,akl->ijkl", inv, inv), dim


def process_weights(povm, input_blochs_T, n_qubits):
    """interval.py:77-88: weights for the Choi Bloch vector; input_blochs_T[s] = bloch(rho_s^T)."""
    povm = np.asarray(povm, dtype=float)
    P, O, D = povm.shape
    flat = povm.reshape(P * O, D)
    S = len(input_blochs_T)
    chan = np.einsum("sd,pi->spdi", np.asarray(input_blochs_T), flat).reshape(S * P * O, D * D)
    dim = 4**n_qubits
    inv = ostate.left_inverse(chan) / dim
    inv = inv.reshape(D * D, S * P, O)
    return np.einsum("aij,akl->ijkl", inv, inv), dim


def quantile(mean, variance, dim, conf_levels, distr_type="gamma", dst="hs"):
    """interval.py:91-110."""
    if distr_type == "norm":
        distr = sts.norm(loc=mean, scale=np.sqrt(variance))
    elif distr_type == "gamma":
        scale = variance / mean
        distr = sts.gamma(a=mean / scale, scale=scale)
    elif distr_type == "exp":
        distr = sts.expon(scale=mean)
    else:
        raise NotImplementedError(f"Unsupported distribution type {distr_type}")
    if dst == "hs":
        alpha = np.sqrt(dim / 2)
    elif dst == "trace":
        alpha = dim / 2
    else:
        raise NotImplementedError()
    return np.sqrt(distr.ppf(np.asarray(conf_levels))) * alpha
