"""Polytope coverage experiments: oracle against the reference's outputs (CPU) and
CUDA kernels against the oracle trial by trial (GPU)."""

import numpy as np
import pytest

from oracle import polytopes as opoly
from oracle.pauli import matrix_to_bloch


def test_oracle_matches_reference_helpers(golden):
    g = golden("polytopes")
    for tag in ("q1", "q2"):
        for i, counts in enumerate(g[f"{tag}_counts"]):
            f = opoly.clipped_frequencies(counts, g[f"{tag}_n_meas"])
            deltas = [opoly.count_delta(cl, f, g[f"{tag}_n_meas"]) for cl in g["levels"]]
            assert np.abs(np.array(deltas) - g[f"{tag}_deltas"][i]).max() < 1e-15
            confs = [opoly.count_confidence(d, f, g[f"{tag}_n_meas"]) for d in (1e-3, 0.02, 0.1)]
            assert np.abs(np.array(confs) - g[f"{tag}_confs"][i]).max() < 1e-15


def test_oracle_trial_is_consistent(golden):
    g = golden("polytopes")
    p_true = opoly.true_probabilities_state(g["q1_povm"], matrix_to_bloch(g["q1_rho"]))
    deltas, inside = opoly.trial_state(g["q1_counts"][0], g["q1_n_meas"], p_true, g["levels"])
    assert np.abs(deltas - g["q1_deltas"][0]).max() < 1e-15
    assert inside[-1] and np.all(np.diff(inside.astype(int)) >= 0)  # larger confidence level -> larger polytope


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["q1", "q2"])
def test_kernel_matches_reference_deltas(golden, tag):
    """Reference counts in -> count_delta / count_confidence within the bisection width (1e-10) / 1e-12."""
    from quantpy_b200.tomography.polytopes import utils

    g = golden("polytopes")
    n_meas = g[f"{tag}_n_meas"]
    f = np.array([opoly.clipped_frequencies(c, n_meas) for c in g[f"{tag}_counts"]])
    got = utils.count_delta_batch(g["levels"], f, n_meas)
    assert np.abs(got - g[f"{tag}_deltas"]).max() < 2e-10
    assert abs(utils.count_delta(0.9, f[0], n_meas) - g[f"{tag}_deltas"][0, 2]) < 2e-10
    for j, d in enumerate((1e-3, 0.02, 0.1)):
        assert abs(utils.count_confidence(d, f[0], n_meas) - g[f"{tag}_confs"][0, j]) < 1e-12


@pytest.mark.gpu
def test_qst_trials_match_oracle_and_reference_coverage(golden):
    import quantpy_b200 as qp
    from quantpy_b200.tomography.polytopes import verification

    g = golden("polytopes")
    state = qp.Qobj(g["qst_state"])
    levels = g["levels"][1:]
    counts, deltas, inside = verification.qst_trials(state, levels, 1000, 2000, seed=5)
    assert counts.shape == (2000, 3, 2) and (counts.sum(-1) == 1000).all()
    povm = qp.generate_measurement_matrix("proj-set", 1)
    p_true = opoly.true_probabilities_state(povm, state.bloch)
    for b in range(0, 2000, 97):
        d, ins = opoly.trial_state(counts[b], np.ones(3) * 1000, p_true, levels)
        assert np.abs(d - deltas[b]).max() < 2e-10
        assert np.array_equal(ins, inside[b])
    cover = inside.mean(0)
    ref = g["qst_cover"]  # 300 reference trials: binomial error bars
    assert np.all(np.abs(cover - ref) < 4 * np.sqrt(ref * (1 - ref) / 300 + 1e-4))
    assert np.all(cover >= levels - 0.02)  # the regions are conservative
    np.random.seed(3)
    again = verification.test_qst(state, levels, 1000, 500)
    assert again.shape == (3,) and np.all(np.abs(again - cover) < 0.06)


@pytest.mark.gpu
def test_qpt_trials_match_oracle_and_reference_coverage(golden):
    import quantpy_b200 as qp
    from oracle import process as oproc
    from quantpy_b200.tomography.polytopes import verification

    g = golden("polytopes")
    chan = qp.channel.depolarizing(0.1, 1)
    levels = g["levels"][1:]
    counts, deltas, inside = verification.qpt_trials(chan, levels, 1000, 1000, seed=8)
    assert counts.shape == (1000, 4, 3, 2)
    povm = qp.generate_measurement_matrix("proj-set", 1)
    choi = oproc.depolarizing_choi(0.1, 1)
    outs = [oproc.apply_choi(choi, r) for r in oproc.input_states("sic", 1)]
    p_true = np.concatenate([opoly.true_probabilities_state(povm, matrix_to_bloch(o)) for o in outs])
    for b in range(0, 1000, 83):
        d, ins = opoly.trial_state(counts[b].reshape(12, 2), np.ones(12) * 1000, p_true, levels, clip_b=False)
        assert np.abs(d - deltas[b]).max() < 2e-10
        assert np.array_equal(ins, inside[b])
    ref = g["qpt_cover"]
    assert np.all(np.abs(inside.mean(0) - ref) < 4 * np.sqrt(ref * (1 - ref) / 100 + 1e-3))
