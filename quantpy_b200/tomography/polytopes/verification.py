"""Coverage experiments for the polytope confidence regions -- drop-ins for test_qst / test_qpt of
quantpy/tomography/polytopes/verification.py:9-78.

The reference simulates `n_trials` tomographies one after another and, for every confidence level, bisects
count_delta and tests whether the true state (channel) lies in the polytope.  Here all trials are sampled by
one launch of the multinomial kernel and evaluated by one launch of k_polytope_coverage (one thread per
trial and level); only the coverage fractions come back to the host.
"""

import ctypes

import numpy as np

from ... import _native as nt
from ... import engine
from ...measurements import generate_measurement_matrix
from ..process import ProcessTomograph
from ..state import StateTomograph


def _coverage(counts, shots, conf_levels, p_true, clip_b):
    torch = nt.torch_cuda()
    lib = nt.load_library()
    B, M, O = counts.shape
    levels = nt.to_device(np.asarray(conf_levels, dtype=np.float64).reshape(-1), torch.float64)
    L = levels.shape[0]
    pt = nt.to_device(np.asarray(p_true, dtype=np.float64).reshape(-1), torch.float64)
    inside = torch.empty((B, L), dtype=torch.uint8, device="cuda")
    delta = torch.empty((B, L), dtype=torch.float64, device="cuda")
    n = np.ascontiguousarray(np.asarray(shots, dtype=np.float64).reshape(-1))
    nt.check(lib.qpb_polytope_coverage(B, M, O, nt.ptr(counts.contiguous()), None, n.ctypes.data_as(ctypes.c_void_p), L,
                                       nt.ptr(levels), nt.ptr(pt), int(clip_b), nt.ptr(delta), nt.ptr(inside),
                                       nt.stream_ptr()))
    return inside, delta


def qst_trials(state, conf_levels, n_measurements=1000, n_trials=1000, seed=None):
    """Per-trial data of test_qst: (counts [B,P,O], deltas [B,L], inside [B,L]) as host arrays."""
    n = state.n_qubits
    povm = generate_measurement_matrix("proj-set", n)  # experiment()'s default POVM, verification.py:14
    P = povm.shape[0]
    shots = np.ones(P) * n_measurements
    dim = 2**n
    flat = povm.reshape(-1, povm.shape[-1])
    # verification.py:17-25: povm0 + (A * dim) . bloch[1:]  (the shot weighting times P cancels for equal shots)
    p_true = flat[:, 0] + (flat[:, 1:] * dim) @ np.asarray(state.bloch)[1:]
    tmg = StateTomograph(state)
    counts = tmg.sample_counts(n_trials, shots, povm, seed=seed, device=True)
    inside, delta = _coverage(counts, shots, conf_levels, p_true, clip_b=True)
    return counts.cpu().numpy(), delta.cpu().numpy(), inside.cpu().numpy().astype(bool)


def test_qst(state, conf_levels, n_measurements=1000, n_trials=1000):
    """Fraction of simulated state tomographies whose polytope at each confidence level contains `state`."""
    _, _, inside = qst_trials(state, conf_levels, n_measurements, n_trials)
    return inside.mean(axis=0)


def qpt_trials(channel, conf_levels, n_measurements=1000, n_trials=1000, input_states="sic", seed=None):
    """Per-trial data of test_qpt: (counts [B,S,P,O], deltas [B,L], inside [B,L])."""
    n = channel.n_qubits
    tmg = ProcessTomograph(channel, input_states=input_states)
    povm = generate_measurement_matrix("proj-set", n)
    P, O = povm.shape[:2]
    shots = np.ones(P) * n_measurements
    # true outcome probabilities of every (input state, POVM, outcome): what verification.py:52-58 assembles as
    # meas0 + A . choi_bloch through the Bloch representation of the Choi matrix
    plan = engine.state_plan(povm, shots)
    outs = [channel.transform(rho) for rho in tmg.input_basis.elements]
    p_true = np.array([(povm.reshape(-1, povm.shape[-1]) @ o.bloch) * 2**n for o in outs]).reshape(-1)
    del plan
    counts = tmg.sample_counts(n_trials, shots, povm, seed=seed, device=True)
    S = counts.shape[1]
    inside, delta = _coverage(counts.reshape(n_trials, S * P, O), np.tile(shots, S), conf_levels, p_true, clip_b=False)
    return counts.cpu().numpy(), delta.cpu().numpy(), inside.cpu().numpy().astype(bool)


def test_qpt(channel, conf_levels, n_measurements=1000, n_trials=1000, input_states="sic"):
    """Fraction of simulated process tomographies whose polytope contains `channel`."""
    _, _, inside = qpt_trials(channel, conf_levels, n_measurements, n_trials, input_states)
    return inside.mean(axis=0)


# these are experiments, not pytest tests, although the reference names them test_* 
test_qst.__test__ = False
test_qpt.__test__ = False
