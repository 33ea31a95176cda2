"""GPU parity tests of the state path: CUDA kernels (through the C ABI / Python shim) against the CPU
oracle and against reference outputs stored in tests/golden.  Tolerances are written in each test:
bit-exact for integer stages, 1e-10 Frobenius for FP64 stages on identical counts."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import distances as odist  # noqa: E402
from oracle import pauli as opauli  # noqa: E402
from oracle import state as ostate  # noqa: E402

STATE_CASES = ["state_c1", "state_c1_pure", "state_c2", "state_c2_set", "state_c2_rank1", "state_c2_sic",
               "state_c3", "state_c3_rank2", "state_c4"]


def fro(a, b):
    return np.sqrt(np.sum(np.abs(np.asarray(a) - np.asarray(b)) ** 2, axis=(-2, -1)))


def haar(n, seed, rank=None):
    rng = np.random.default_rng(seed)
    d = 2**n
    k = d if rank is None else rank
    g = rng.normal(size=(d, k)) + 1j * rng.normal(size=(d, k))
    rho = g @ g.conj().T
    return rho / np.trace(rho)


@pytest.fixture(scope="module")
def qp():
    import quantpy_b200

    return quantpy_b200


def tomograph(qp, g):
    tmg = qp.StateTomograph(qp.Qobj(g["rho_true"]))
    tmg.povm_matrix = g["povm_matrix"]
    tmg.results = g["counts"][0]
    tmg.n_measurements = g["n_meas"]
    return tmg


# ----------------------------------------------------------------------------- probabilities, sampler

@pytest.mark.parametrize("n,povm,B", [(2, "proj", 5000), (3, "proj", 3000), (3, "sic", 1000 + 11), (4, "proj", 300),
                                      (2, "proj-set", 63)])
def test_batched_probabilities_on_dmma_match_oracle(qp, n, povm, B):
    """qpb_povm_probs for thousands of states is one dense contraction p = 2^n M r (state.py:109-110): it runs on the
    FP64 tensor cores (TMA-staged DMMA GEMM) for B >= 64 and on the warp-reduction kernel below; both within 1e-13
    of the oracle, clipped to [0, 1]."""
    from quantpy_b200 import _native as nt
    from quantpy_b200 import engine

    pm = qp.generate_measurement_matrix(povm, n)
    plan = engine.state_plan(pm, np.ones(pm.shape[0]) * 10000)
    rng = np.random.default_rng(n)
    blochs = np.array([opauli.matrix_to_bloch(haar(n, 100 + i, rank=1 + i % 2**n)) for i in range(24)])
    blochs = blochs[rng.integers(0, 24, B)] * (1 + 1e-3 * rng.normal(size=(B, 1)))
    want = np.clip(np.einsum("ijk,bk->bij", pm, blochs) * 2**n, 0, 1).reshape(B, -1)
    got = plan.probabilities(blochs).cpu().numpy()
    assert np.abs(got - want).max() < 1e-13
    with nt.option("NO_TMA_GEMM", 1):
        ref = plan.probabilities(blochs).cpu().numpy()
    assert np.abs(ref - want).max() < 1e-13

@pytest.mark.parametrize("case", STATE_CASES)
def test_probabilities_match_reference(qp, golden, case):
    from quantpy_b200 import engine

    g = golden(case)
    plan = engine.state_plan(g["povm_matrix"], g["n_meas"])
    p = plan.probabilities(opauli.matrix_to_bloch(g["rho_true"])).cpu().numpy().reshape(g["probs"].shape)
    assert np.abs(p - g["probs"]).max() < 1e-14


@pytest.fixture(params=["binomial", "alias", "binomial-1-group", "binomial-3-groups", "binomial-8-groups"])
def sampler_kind(request):
    """Both multinomial kernels -- and the conditional-binomial one as a single chain and with its outcomes in 3 / 8
    groups (SAMPLER_LANES; the default picks the number of groups from the number of outcomes: a pass that splits the
    shots between the groups, then a chain per group) -- are held to the same distributional tests."""
    from quantpy_b200 import _native as nt

    kind, _, lanes = request.param.partition("-")
    with nt.option("SAMPLER", nt.SAMPLERS[kind]), nt.option("SAMPLER_LANES", int(lanes[0]) if lanes else 0):
        yield request.param


@pytest.fixture
def forced_binomial():
    from quantpy_b200 import _native as nt

    with nt.option("SAMPLER", nt.SAMPLERS["binomial"]):
        yield


@pytest.mark.parametrize("n_shots,p", [(10000, 0.3), (10000, 1 / 36), (10000, 0.001), (10000, 0.9995), (100, 0.5),
                                       (57, 0.93), (1000000, 0.41), (3, 0.2), (2000, 0.015), (500, 0.06)])
def test_binomial_kernel_matches_exact_pmf(qp, forced_binomial, n_shots, p):
    """O = 2 makes the multinomial a single binomial draw: chi-square against scipy.stats.binom over
    both regimes (inversion for n*min(p,1-p) < 30, BTPE otherwise) and the p > 1/2 reflection."""
    import torch
    from scipy import stats

    from quantpy_b200 import engine

    B = 400000
    probs = torch.tensor([p, 1 - p], dtype=torch.float64, device="cuda")
    counts = engine.sample_counts(probs, B, 1, 2, [n_shots], seed=7 + n_shots, offset=0).cpu().numpy()[:, 0, :]
    assert (counts.sum(-1) == n_shots).all()
    x = counts[:, 0]
    assert abs(x.mean() - n_shots * p) < 6 * np.sqrt(n_shots * p * (1 - p) / B)
    assert abs(x.var() / (n_shots * p * (1 - p)) - 1) < 0.02
    lo, hi = int(x.min()), int(x.max())
    ks = np.arange(lo, hi + 1)
    observed = np.bincount(x - lo, minlength=len(ks)).astype(float)
    expected = stats.binom.pmf(ks, n_shots, p) * B
    # merge the sparse tails into their neighbours so that every cell expects >= 10 draws
    keep = expected >= 10
    first, last = np.argmax(keep), len(keep) - 1 - np.argmax(keep[::-1])
    obs = np.r_[observed[:first + 1].sum(), observed[first + 1:last], observed[last:].sum()]
    exp = np.r_[stats.binom.cdf(ks[first], n_shots, p) * B, expected[first + 1:last],
                stats.binom.sf(ks[last] - 1, n_shots, p) * B]
    stat = ((obs - exp) ** 2 / exp).sum()
    assert stats.chi2.sf(stat, len(obs) - 1) > 1e-4, (stat, len(obs))


def test_binomial_prefilter_changes_no_count(qp, forced_binomial):
    """The float32 prefilter of the BTRS acceptance test only settles candidates whose float64 verdict is certain,
    so the counts are the same integers with and without it -- across shot numbers, skewed and flat tables, block
    sizes and test periods (a lane's stream is its own)."""
    import torch

    from quantpy_b200 import _native as nt
    from quantpy_b200 import engine

    rng = np.random.default_rng(5)
    cases = [(36, 10000, 20000), (36, 1000000, 4000), (8, 300, 20000), (216, 100000, 2000), (2, 50000000, 20000)]
    for O, shots, B in cases:
        p = rng.dirichlet(np.full(O, 0.7))
        probs = torch.tensor(p[None], dtype=torch.float64, device="cuda")
        with nt.option("SAMPLER_NO_PREFILTER", 1):
            want = engine.sample_counts(probs, B, 1, O, [shots], seed=11 + O, offset=3).cpu().numpy()
        got = engine.sample_counts(probs, B, 1, O, [shots], seed=11 + O, offset=3).cpu().numpy()
        assert (got.sum(-1) == shots).all()
        assert np.array_equal(got, want), (O, shots)
        with nt.option("SAMPLER_EXACT_EVERY", 1), nt.option("SAMPLER_THREADS", 96):
            other = engine.sample_counts(probs, B, 1, O, [shots], seed=11 + O, offset=3).cpu().numpy()
        assert np.array_equal(other, want), (O, shots)


@pytest.mark.parametrize("n,povm", [(1, "proj-set"), (2, "proj"), (2, "proj-set"), (3, "proj"), (4, "proj")])
def test_sampler_counts_sum_and_chi_square(qp, sampler_kind, n, povm):
    """Integer stage: every POVM's counts sum to the shot number exactly.  Distribution: chi-square per
    outcome against the exact probabilities (NumPy's stream is not reproduced)."""
    from scipy import stats

    rho = haar(n, 10 + n)
    tmg = qp.StateTomograph(qp.Qobj(rho))
    shots = 10000
    B = 2000 if n < 4 else 200
    counts = tmg.sample_counts(B, shots, povm, seed=123)
    povm_matrix = qp.generate_measurement_matrix(povm, n)
    assert counts.shape == (B,) + povm_matrix.shape[:2]
    assert counts.dtype == np.int64
    assert np.array_equal(counts.sum(-1), np.full(counts.shape[:2], shots))
    probs = ostate.probabilities(povm_matrix, opauli.matrix_to_bloch(rho))
    # pooled chi-square per POVM (B*shots draws), then the per-sample statistic's mean
    for m in range(povm_matrix.shape[0]):
        pooled = counts[:, m].sum(0)
        expect = probs[m] / probs[m].sum() * pooled.sum()
        keep = expect > 20
        stat = ((pooled[keep] - expect[keep]) ** 2 / expect[keep]).sum()
        dof = keep.sum() - 1
        assert stats.chi2.sf(stat, dof) > 1e-4, (m, stat, dof)
    # variance check on one outcome: binomial variance n p (1-p)
    k = int(np.argmax(probs[0]))
    pk = probs[0, k] / probs[0].sum()
    var = counts[:, 0, k].var()
    assert abs(var / (shots * pk * (1 - pk)) - 1) < (0.15 if n < 4 else 0.4)


def test_multinomial_covariance(qp, sampler_kind):
    """Joint structure: Cov(n_i, n_j) = -N p_i p_j for the conditional-binomial chain and the alias sampler."""
    import torch

    from quantpy_b200 import engine

    p = np.array([0.5, 0.2, 0.15, 0.1, 0.05])
    N, B = 5000, 200000
    counts = engine.sample_counts(torch.tensor(p, device="cuda"), B, 1, 5, [N], seed=3).cpu().numpy()[:, 0, :].astype(float)
    cov = np.cov(counts.T)
    want = N * (np.diag(p) - np.outer(p, p))
    assert np.abs(cov - want).max() < 0.03 * N * p.max()
    assert np.abs(counts.mean(0) - N * p).max() < 0.2


def test_sampler_is_shard_invariant_and_seeded(qp, sampler_kind):
    """The draw for a global sample index does not depend on the batch split (multi-GPU sharding)."""
    tmg = qp.StateTomograph(qp.Qobj(haar(2, 3)))
    full = tmg.sample_counts(64, 10000, "proj", seed=99)
    a = tmg.sample_counts(40, 10000, "proj", seed=99, offset=0)
    b = tmg.sample_counts(24, 10000, "proj", seed=99, offset=40)
    assert np.array_equal(full, np.concatenate([a, b]))
    other = tmg.sample_counts(64, 10000, "proj", seed=100)
    assert not np.array_equal(full, other)
    np.random.seed(5)
    x = tmg.sample_counts(4, 10000, "proj")
    np.random.seed(5)
    y = tmg.sample_counts(4, 10000, "proj")
    assert np.array_equal(x, y)


def test_sampler_edge_cases(qp, sampler_kind):
    # a pure |0> state: z- outcome has probability exactly 0 and must never appear
    tmg = qp.StateTomograph(qp.Qobj([1, 0], is_ket=True))
    counts = tmg.sample_counts(500, [7, 1, 12345], "proj-set", seed=1)
    assert np.array_equal(counts.sum(-1), np.tile([7, 1, 12345], (500, 1)))
    assert counts[:, 2, 1].max() == 0 and (counts[:, 2, 0] == 12345).all()
    assert tmg.sample_counts(0, 10, "proj-set", seed=1).shape == (0, 3, 2)
    with pytest.raises(ValueError):
        tmg.sample_counts(3, [1, 2], "proj-set")


# ----------------------------------------------------------------------------- linear inversion

@pytest.mark.parametrize("case", STATE_CASES)
def test_lin_matches_reference_outputs(qp, golden, case):
    """Reference counts in -> rho within 1e-10 (Frobenius) of the reference's own output."""
    g = golden(case)
    tmg = tomograph(qp, g)
    for physical, key in ((True, "lin_physical"), (False, "lin_raw")):
        got = tmg.point_estimate_batch(g["counts"], "lin", physical=physical)
        assert fro(got, g[key]).max() < 1e-10
    one = tmg.point_estimate("lin")
    assert fro(one.matrix, g["lin_physical"][0]) < 1e-10
    assert fro(tmg.reconstructed_state.matrix, one.matrix) == 0


@pytest.mark.parametrize("n,B", [(1, 3000), (2, 3000), (3, 600), (4, 40)])
def test_lin_matches_oracle_on_gpu_counts(qp, n, B):
    rho = haar(n, 20 + n, rank=None if n != 3 else 2)
    tmg = qp.StateTomograph(qp.Qobj(rho))
    np.random.seed(n)
    tmg.experiment(10000, "proj")
    counts = tmg.sample_counts(B, 10000, "proj", seed=n)
    got = tmg.point_estimate_batch(counts, "lin")
    want = ostate.lin_estimate(counts, tmg.povm_matrix, tmg.n_measurements)
    assert fro(got, want).max() < 1e-10
    assert np.abs(np.trace(got, axis1=1, axis2=2) - 1).max() < 1e-12
    assert np.linalg.eigvalsh(got).min() > 0


@pytest.mark.parametrize("n", [3, 4])
def test_projection_on_degenerate_and_extreme_spectra(qp, n):
    """The register-resident Jacobi (one lane per matrix row) on inputs that stress it: a maximally degenerate
    spectrum (uniform counts -> identity / d, already diagonal), counts of a pure product state (rank 1, most
    eigenvalues clipped at 1e-15), and a batch size that leaves lane groups of the last warp without a matrix."""
    povm = qp.generate_measurement_matrix("proj", n)
    K = povm.shape[1]
    tmg = qp.StateTomograph(qp.Qobj(np.eye(2**n) / 2**n))
    tmg.povm_matrix = povm
    tmg.n_measurements = np.ones(1) * 6**n * 10
    uniform = np.full((1, K), 10, dtype=np.int64)
    ket = np.zeros(2**n); ket[0] = 1.0
    pure = np.outer(ket, ket)
    p = np.clip(povm[0] @ qp.Qobj(pure).bloch * 2**n, 0, 1)
    exact = np.rint(p * 6**n * 10).astype(np.int64)[None]
    exact[0, np.argmax(exact[0])] += 6**n * 10 - exact.sum()
    rng = np.random.default_rng(n)
    noisy = rng.multinomial(6**n * 10, p, size=3)
    counts = np.concatenate([uniform, exact, noisy])  # B = 5
    got = tmg.point_estimate_batch(counts[:, None, :], "lin")
    want = ostate.lin_estimate(counts[:, None, :], povm, tmg.n_measurements)
    assert fro(got, want).max() < 1e-10
    assert fro(got[0], np.eye(2**n) / 2**n) < 1e-12
    assert np.abs(np.trace(got, axis1=1, axis2=2) - 1).max() < 1e-12
    assert np.linalg.eigvalsh(got).min() > -1e-15


# ----------------------------------------------------------------------------- maximum likelihood

@pytest.mark.parametrize("case", ["state_c1", "state_c1_pure", "state_c2", "state_c2_set", "state_c2_rank1",
                                  "state_c2_sic", "state_c3", "state_c3_rank2"])
@pytest.mark.parametrize("init", ["lin", "mixed"])
def test_mle_fixed_iterations_match_oracle(qp, golden, case, init):
    """Same update, same start, same number of iterations -> 1e-10."""
    g = golden(case)
    tmg = tomograph(qp, g)
    iters = 40 if g["rho_true"].shape[0] <= 4 else 12
    got, its = tmg.point_estimate_batch(g["counts"], "mle", init=init, max_iter=iters, tol=0.0, return_iters=True)
    want = ostate.mle_rrr(g["counts"], g["povm_matrix"], g["n_meas"], init=init, max_iter=iters, tol=0.0)
    assert (its == iters).all()
    assert fro(got, want).max() < 1e-10


def test_mle_four_qubits_matches_oracle(qp, golden):
    g = golden("state_c4")
    tmg = tomograph(qp, g)
    got = tmg.point_estimate_batch(g["counts"], "mle", max_iter=3, tol=0.0)
    want = ostate.mle_rrr(g["counts"], g["povm_matrix"], g["n_meas"], max_iter=3, tol=0.0)
    assert fro(got, want).max() < 1e-10


@pytest.mark.parametrize("n,povm,B", [(1, "proj-set", 2000), (2, "proj", 2000), (2, "proj-set", 500)])
@pytest.mark.parametrize("tol,max_iter", [(1e-3, 100), (1e-6, 300)])
def test_mle_tolerance_mode_matches_oracle(qp, n, povm, B, tol, max_iter):
    """Per-sample stopping: iteration counts equal the oracle's and states agree to 1e-10."""
    rho = haar(n, 40 + n)
    tmg = qp.StateTomograph(qp.Qobj(rho))
    np.random.seed(n)
    tmg.experiment(10000, povm)
    counts = tmg.sample_counts(B, 10000, povm, seed=5)
    got, its = tmg.point_estimate_batch(counts, "mle", max_iter=max_iter, tol=tol, return_iters=True)
    want, wits = ostate.mle_rrr(counts, tmg.povm_matrix, tmg.n_measurements, max_iter=max_iter, tol=tol,
                                return_iters=True)
    same = its == wits
    assert same.mean() > 0.999  # a step norm within rounding of tol may stop one iteration apart
    assert fro(got[same], want[same]).max() < 1e-10
    assert 1 <= its.mean() < max_iter
    _check_off_by_one(tmg, counts, got, its, wits)


def _check_off_by_one(tmg, counts, got, its, wits):
    """Samples whose stopping iteration differs from the oracle's (step norm within rounding of tol) must differ by
    exactly one iteration and equal the oracle's iterate at the kernel's own count to 1e-10."""
    for i in np.flatnonzero(its != wits):
        assert abs(int(its[i]) - int(wits[i])) == 1, (i, its[i], wits[i])
        step = ostate.mle_rrr(counts[i: i + 1], tmg.povm_matrix, tmg.n_measurements, max_iter=int(its[i]), tol=0.0)
        assert fro(got[i: i + 1], step).max() < 1e-10


def test_mle_bench_setting_stragglers_match_oracle(qp):
    """The exact bench.py setting (BASELINE configs[1]: 2 qubits, 'proj', 1e4 shots, B = 1e5, tol 1e-6, max_iter 1000),
    where long-running samples change lane mapping mid-flight: the 200 longest-running samples and a 1 % stride
    are re-derived by the oracle -- same iteration counts (or one apart, checked at one-step distance), states 1e-10."""
    rho = haar(2, 0)
    tmg = qp.StateTomograph(qp.Qobj(rho))
    np.random.seed(0)
    tmg.experiment(10000, "proj")
    B = 100000
    counts = tmg.sample_counts(B, 10000, "proj", seed=1234)
    got, its = tmg.point_estimate_batch(counts, "mle", max_iter=1000, tol=1e-6, return_iters=True)
    assert its.max() > 600 and its.mean() > 100
    pick = np.unique(np.concatenate([np.argsort(-its)[:200], np.arange(0, B, 100)]))
    want, wits = ostate.mle_rrr(counts[pick], tmg.povm_matrix, tmg.n_measurements, max_iter=1000, tol=1e-6,
                                return_iters=True)
    same = its[pick] == wits
    assert same.mean() > 0.995
    assert fro(got[pick][same], want[same]).max() < 1e-10
    _check_off_by_one(tmg, counts[pick], got[pick], its[pick], wits)


@pytest.mark.parametrize("n,B,max_iter", [(3, 331, 300), (4, 70, 12)])
def test_tiled_general_povm_mle_matches_oracle(qp, n, B, max_iter):
    """k_gemm_counts_dmma-based batched R.rho.R ('sic' POVM: no Pauli-axis structure) with per-sample stopping,
    compaction of the running set every 8 iterations and a ragged batch: iteration counts equal the oracle's
    (or one apart, checked at one-step distance), states 1e-10; the warp-per-sample kernel gives the same."""
    from quantpy_b200 import _native as nt
    from quantpy_b200 import engine

    rho = haar(n, 70 + n)
    tmg = qp.StateTomograph(qp.Qobj(rho))
    np.random.seed(n)
    tmg.experiment(10000, "sic")
    plan = engine.state_plan(tmg.povm_matrix, tmg.n_measurements)
    assert nt.load_library().qpb_mle_variant(plan.handle) == 5
    counts = tmg.sample_counts(B, 10000, "sic", seed=9)
    tol = 1e-5 if n == 3 else 0.0
    got, its = tmg.point_estimate_batch(counts, "mle", max_iter=max_iter, tol=tol, return_iters=True)
    want, wits = ostate.mle_rrr(counts, tmg.povm_matrix, tmg.n_measurements, max_iter=max_iter, tol=tol,
                                return_iters=True)
    same = its == wits
    assert same.mean() > 0.99
    assert fro(got[same], want[same]).max() < 1e-10
    _check_off_by_one(tmg, counts, got, its, wits)
    if n == 3:
        assert its.min() < its.max() and 1 < its.mean() < max_iter
    with nt.option("NO_TILED_MLE", 1):
        ref, rits = tmg.point_estimate_batch(counts, "mle", max_iter=max_iter, tol=tol, return_iters=True)
    assert (rits == its).mean() > 0.99
    assert fro(got[rits == its], ref[rits == its]).max() < 1e-11
    mixed = tmg.point_estimate_batch(counts[:40], "mle", init="mixed", max_iter=3, tol=0.0)
    assert fro(mixed, ostate.mle_rrr(counts[:40], tmg.povm_matrix, tmg.n_measurements, init="mixed", max_iter=3,
                                     tol=0.0)).max() < 1e-10


def test_mle_four_qubits_deep_iterations_match_oracle(qp):
    """k_mle_rrr_axis<4> at the depth the bench runs it: 200 iterations on 32 count tables, a quarter of them
    drawn from a rank-1 state (zero counts, probabilities at the 1e-10 guard) -- DMMA accumulation order and
    the axis maps stay within 1e-10 of the dense oracle."""
    full, pure = haar(4, 11), haar(4, 12, rank=1)
    povm = qp.generate_measurement_matrix("proj", 4)
    tables = []
    for state, m, seed in ((full, 24, 1), (pure, 8, 2)):
        tmg = qp.StateTomograph(qp.Qobj(state))
        np.random.seed(seed)
        tmg.experiment(10000, "proj")
        tables.append(tmg.sample_counts(m, 10000, "proj", seed=seed))
    counts = np.concatenate(tables)
    got, its = tmg.point_estimate_batch(counts, "mle", max_iter=200, tol=0.0, return_iters=True)
    want = ostate.mle_rrr(counts, povm, np.ones(1) * 10000, max_iter=200, tol=0.0)
    assert (its == 200).all()
    assert fro(got, want).max() < 1e-10


@pytest.mark.parametrize("n,povm,disable,expect", [
    (2, "proj", [], 3), (2, "proj", ["PAULI"], 2), (2, "proj", ["PAULI", "CONST"], 1),
    (2, "proj-set", [], 3), (2, "proj4", [], 3), (2, "sic", [], 2), (2, "sic", ["CONST"], 1),
    (1, "proj-set", [], 2), (1, "proj-set", ["CONST"], 1),
    (3, "proj", [], 4), (3, "proj", ["AXIS"], 5), (3, "proj", ["AXIS", "TILED_MLE"], 0), (3, "proj-set", [], 4), (3, "sic", [], 5), (3, "sic", ["TILED_MLE"], 0), (4, "proj", [], 4), (4, "sic", [], 5),
])
def test_every_mle_kernel_variant_matches_oracle(qp, n, povm, disable, expect):
    """qpb_mle_rrr dispatches on the POVM's structure; every variant is the same update to 1e-10."""
    import contextlib

    from quantpy_b200 import _native as nt
    from quantpy_b200 import engine

    with contextlib.ExitStack() as stack:
        for name in disable:
            stack.enter_context(nt.option(f"NO_{name}" if name == "TILED_MLE" else f"NO_{name}_KERNEL", 1))
        _check_mle_variant(qp, n, povm, expect)


def _check_mle_variant(qp, n, povm, expect):
    from quantpy_b200 import _native as nt
    from quantpy_b200 import engine

    rho = haar(n, 90 + n, rank=2 if n == 3 else None)
    tmg = qp.StateTomograph(qp.Qobj(rho))
    np.random.seed(n)
    tmg.experiment(10000, povm)
    plan = engine.state_plan(tmg.povm_matrix, tmg.n_measurements)
    assert nt.load_library().qpb_mle_variant(plan.handle) == expect
    B = {1: 500, 2: 500, 3: 60, 4: 6 if expect != 5 else 40}[n]
    counts = tmg.sample_counts(B, 10000, povm, seed=17)
    tol, max_iter = (1e-5, 40) if n < 4 else (0.0, 4)
    got, its = tmg.point_estimate_batch(counts, "mle", max_iter=max_iter, tol=tol, return_iters=True)
    want, wits = ostate.mle_rrr(counts, tmg.povm_matrix, tmg.n_measurements, max_iter=max_iter, tol=tol,
                                return_iters=True)
    assert np.array_equal(its, wits)
    assert fro(got, want).max() < 1e-10
    mixed = tmg.point_estimate_batch(counts[:4], "mle", init="mixed", max_iter=5, tol=0.0)
    assert fro(mixed, ostate.mle_rrr(counts[:4], tmg.povm_matrix, tmg.n_measurements, init="mixed", max_iter=5,
                                     tol=0.0)).max() < 1e-10


def test_mle_with_unequal_shots_per_povm(qp):
    """Shot weights enter the POVM table (state.py:193-196); the structured kernels carry them as per-slot guards."""
    for n in (2, 3):
        rho = haar(n, 7)
        tmg = qp.StateTomograph(qp.Qobj(rho))
        shots = (np.arange(3**n) % 4 + 1) * 2500
        np.random.seed(1)
        tmg.experiment(shots, "proj-set")
        assert np.array_equal(tmg.results.sum(-1), shots)
        counts = tmg.sample_counts(20, shots, "proj-set", seed=2)
        got = tmg.point_estimate_batch(counts, "mle", max_iter=30, tol=0.0)
        want = ostate.mle_rrr(counts, tmg.povm_matrix, tmg.n_measurements, max_iter=30, tol=0.0)
        assert fro(got, want).max() < 1e-10
        lin = tmg.point_estimate_batch(counts, "lin")
        assert fro(lin, ostate.lin_estimate(counts, tmg.povm_matrix, tmg.n_measurements)).max() < 1e-10


@pytest.mark.parametrize("case", ["state_c1", "state_c2", "state_c2_set", "state_c2_rank1", "state_c2_sic"])
def test_mle_is_at_least_as_likely_as_reference(qp, golden, case):
    """Loose pin to the reference's BFGS 'mle' (SURVEY D1): our likelihood is never worse."""
    g = golden(case)
    tmg = tomograph(qp, g)
    got = tmg.point_estimate_batch(g["counts"], "mle", max_iter=20000, tol=1e-13)
    for i, c in enumerate(g["counts"]):
        ours = ostate.neg_log_likelihood(got[i], c, g["povm_matrix"], g["n_meas"])
        ref = ostate.neg_log_likelihood(g["mle_default"][i], c, g["povm_matrix"], g["n_meas"])
        assert ours <= ref + 1e-9


def test_mle_edge_cases(qp):
    tmg = qp.StateTomograph(qp.Qobj([1, 0], is_ket=True))
    np.random.seed(0)
    tmg.experiment(10000, "proj-set")  # contains an outcome with zero counts
    assert tmg.results[2, 1] == 0
    counts = np.repeat(tmg.results[None], 5, axis=0)
    got, its = tmg.point_estimate_batch(counts, "mle", max_iter=500, tol=1e-9, return_iters=True)
    want, wits = ostate.mle_rrr(counts, tmg.povm_matrix, tmg.n_measurements, max_iter=500, tol=1e-9,
                                return_iters=True)
    assert np.isfinite(got).all() and np.array_equal(its, wits)
    assert fro(got, want).max() < 1e-10
    # max_iter = 0 returns the start state; empty batch is allowed
    start = tmg.point_estimate_batch(counts, "lin")
    assert fro(tmg.point_estimate_batch(counts, "mle", max_iter=0), start).max() < 1e-15
    assert tmg.point_estimate_batch(counts[:0], "mle").shape == (0, 2, 2)
    with pytest.raises(ValueError):
        tmg.point_estimate("nope")
    with pytest.raises(ValueError):
        tmg.point_estimate("mle", init="nope")


@pytest.mark.parametrize("case", ["state_c1", "state_c2_set"])
def test_mle_constr_runs_the_same_kernel_and_dominates_reference(qp, golden, case):
    """'mle-constr' (reference: SLSQP under Tr rho = 1, state.py:231-254) maximises the same likelihood over the
    same set; here it is the R.rho.R kernel.  Tier-2 parity as for 'mle': at least as likely as the reference's
    answer, and equal to it within the reference optimiser's own noise floor."""
    g, ref = golden(case), golden("mle_constr")
    tmg = qp.StateTomograph(qp.Qobj(g["rho_true"]))
    tmg.povm_matrix = g["povm_matrix"]
    for i, c in enumerate(g["counts"][:4]):
        tmg.results = c
        tmg.n_measurements = g["n_meas"]
        a = tmg.point_estimate("mle-constr", tol=1e-12, max_iter=20000).matrix
        b = tmg.point_estimate("mle", tol=1e-12, max_iter=20000).matrix
        assert np.array_equal(a, b)
        ours = ostate.neg_log_likelihood(a, c, g["povm_matrix"], g["n_meas"])
        for key in ("_default", "_tight"):
            theirs = ostate.neg_log_likelihood(ref[case + key][i], c, g["povm_matrix"], g["n_meas"])
            assert ours <= theirs + 1e-9
        assert np.linalg.norm(a - ref[case + "_tight"][i]) < 2e-6


# ----------------------------------------------------------------------------- distances

@pytest.mark.parametrize("case", STATE_CASES)
def test_distances_match_reference(qp, golden, case):
    from quantpy_b200 import _native as nt
    from quantpy_b200 import engine

    g = golden(case)
    est = nt.complex_to_device(g["lin_physical"])
    hs = engine.distance(est, g["rho_true"], "hs").cpu().numpy()
    tr = engine.distance(est, g["rho_true"], "trace").cpu().numpy()
    inf = engine.distance(est, g["rho_true"], "if").cpu().numpy()
    assert np.abs(hs - g["dist_hs"]).max() < 1e-13
    # vs the reference's sqrtm-based values: 1e-10 where the spectra are regular, 1e-6 where sqrtm
    # itself is noisy (see tests/test_oracle_golden.py::sqrtm_tolerances)
    from test_oracle_golden import sqrtm_tolerances

    tol_tr, tol_if = sqrtm_tolerances(g["lin_physical"], g["rho_true"])
    assert (np.abs(tr - g["dist_trace"]) < tol_tr).all()
    assert (np.abs(inf - g["dist_if"]) < tol_if).all()
    # vs the oracle's eigenvalue forms (same algorithm): tight for hs/trace; the infidelity takes a
    # square root of eigenvalues that may be ~1e-15, so rounding moves it by ~sqrt(1e-16)
    assert np.abs(tr - odist.trace(g["lin_physical"], g["rho_true"])).max() < 1e-13
    assert (np.abs(inf - odist.infidelity(g["lin_physical"], g["rho_true"])) < np.where(tol_if < 1e-9, 1e-10, 1e-6)).all()
    same = engine.distance(est[:1], g["lin_physical"][0], "hs").cpu().numpy()
    assert same[0] == 0.0


# ----------------------------------------------------------------------------- single-experiment API

def test_state_tomograph_api(qp):
    rho = haar(2, 77)
    tmg = qp.StateTomograph(qp.Qobj(rho), dst="trace")
    assert tmg.dst is qp.trace_dst
    np.random.seed(1)
    tmg.experiment(1000)  # default 'proj-set'
    assert tmg.povm_matrix.shape == (9, 4, 16)
    assert tmg.results.shape == (9, 4) and tmg.results.dtype == np.int64
    assert np.array_equal(tmg.n_measurements, np.full(9, 1000.0))
    assert tmg.flat_results.shape == (36,)
    first = tmg.results.copy()
    tmg.experiment(500, warm_start=True)
    assert tmg.results.shape == (18, 4) and np.array_equal(tmg.results[:9], first)
    assert np.array_equal(tmg.n_measurements, np.r_[np.full(9, 1000), np.full(9, 500)])
    est = tmg.point_estimate("lin")
    want = ostate.lin_estimate(tmg.results, tmg.povm_matrix, tmg.n_measurements)
    assert fro(est.matrix, want) < 1e-10
    # results setter recomputes n_measurements exactly (integer stage)
    tmg.results = first
    assert np.array_equal(tmg.n_measurements, first.sum(-1)) and tmg.n_measurements.dtype == np.int64
    with pytest.raises(ValueError):
        tmg.experiment([10, 20], "proj-set")
    with pytest.raises(ValueError):
        qp.StateTomograph(qp.Qobj(rho), dst="nope")


def _pauli2_counts(qp, B, povm="proj", state_seed=5, seed=3):
    rho = haar(2, state_seed)
    tmg = qp.StateTomograph(qp.Qobj(rho))
    np.random.seed(0)
    tmg.experiment(10000, povm)
    return tmg, tmg.sample_counts(B, 10000, povm, seed=seed)


@pytest.mark.parametrize("povm", ["proj", "proj-set"])
def test_pauli2_lane_mappings_are_bit_identical(qp, povm):
    """k_mle_rrr_pauli2 evaluates ONE dataflow graph with two lane mappings (thread per sample in registers, warp
    per sample with the products on DMMA).  Thread-only, warp-only and the production mix (hand-over of long-running
    samples at several ages, with and without dedicated workers) must agree in every bit of every state and in
    every iteration count; 'proj-set' has non-uniform guards (the second template instance)."""
    from quantpy_b200 import _native as nt

    tmg, counts = _pauli2_counts(qp, 40000, povm)
    kw = dict(max_iter=400, tol=1e-6, return_iters=True)
    with nt.option("MLE_LANES", 1):
        a, ia = tmg.point_estimate_batch(counts, "mle", **kw)
    with nt.option("MLE_LANES", 32):
        b, ib = tmg.point_estimate_batch(counts, "mle", **kw)
    assert np.array_equal(ia, ib) and np.array_equal(a, b)
    assert ia.max() == 400 and ia.min() < 50
    for age, live, workers in [(0, 0, 0), (1, 32, 4), (37, 3, -1), (150, 12, 2)]:
        with nt.option("MLE_PARK_AGE", age), nt.option("MLE_PARK_LIVE", live), nt.option("MLE_W_WARPS", workers):
            c, ic = tmg.point_estimate_batch(counts, "mle", **kw)
        assert np.array_equal(ia, ic) and np.array_equal(a, c), (age, live, workers)
    # init='mixed' (no start state) and a batch small enough to take the warp-only path on its own
    with nt.option("MLE_LANES", 1):
        a, ia = tmg.point_estimate_batch(counts[:3000], "mle", init="mixed", **kw)
    c, ic = tmg.point_estimate_batch(counts[:3000], "mle", init="mixed", **kw)
    assert np.array_equal(ia, ic) and np.array_equal(a, c)


def test_pauli2_fused_distance_is_bit_identical(qp):
    """The bootstrap's Hilbert-Schmidt distance comes out of the MLE kernel's write-back (both lane mappings) or out
    of k_distance (NO_HS_FUSION): same formula, same `< 1e-15 -> 0` rule (quantpy/geometry.py:17-18), same bits."""
    import torch

    from quantpy_b200 import _native as nt
    from quantpy_b200 import engine

    rho = haar(2, 5)
    state = qp.Qobj(rho)
    povm = qp.generate_measurement_matrix("proj", 2)
    plan = engine.state_plan(povm, np.ones(1) * 10000)
    probs = plan.probabilities(state.bloch)[0].contiguous()
    kw = dict(method="mle", max_iter=300, tol=1e-6, dst="hs")
    out = {}
    for name, opts in {"fused": {}, "thread": {"MLE_LANES": 1}, "warp": {"MLE_LANES": 32}, "unfused": {"NO_HS_FUSION": 1}}.items():
        import contextlib

        with contextlib.ExitStack() as stack:
            for k, v in opts.items():
                stack.enter_context(nt.option(k, v))
            got = plan.bootstrap(probs, 20000, 77, 0, rho, **kw)
            torch.cuda.synchronize()
            out[name] = (got["dist"].cpu().numpy(), got["iters"].cpu().numpy())
    for name in ("thread", "warp", "unfused"):
        assert np.array_equal(out["fused"][1], out[name][1]), name
        assert np.array_equal(out["fused"][0], out[name][0]), name


def test_tail_merging_is_bit_identical(qp):
    """The structured kernel hands a warp's last samples to its partner warp once the queue is empty; every
    sample still runs its own iteration sequence, so the output must not change by a single bit."""
    rho = haar(2, 5)
    tmg = qp.StateTomograph(qp.Qobj(rho))
    np.random.seed(0)
    tmg.experiment(10000, "proj")
    counts = tmg.sample_counts(60000, 10000, "proj", seed=3)
    from quantpy_b200 import _native as nt

    a, ia = tmg.point_estimate_batch(counts, "mle", max_iter=400, tol=1e-6, return_iters=True)
    with nt.option("NO_TAIL_MERGE", 1):
        b, ib = tmg.point_estimate_batch(counts, "mle", max_iter=400, tol=1e-6, return_iters=True)
    assert np.array_equal(ia, ib) and np.array_equal(a, b)
    assert ia.max() == 400 and ia.min() < 50


def test_ordered_entry_points_change_no_bit(qp):
    """qpb_lin_project_ordered returns the physical estimates of qpb_lin_project plus a permutation that puts the
    samples with the smallest positive eigenvalue of the unprojected estimate first; qpb_mle_rrr_ordered gives
    qpb_mle_rrr's bits."""
    import torch

    from quantpy_b200 import _native as nt
    from quantpy_b200 import engine

    lib = nt.load_library()
    rho = haar(2, 0)
    povm = qp.generate_measurement_matrix("proj", 2)
    plan = engine.state_plan(povm, np.ones(1) * 10000)
    probs = plan.probabilities(qp.Qobj(rho).bloch)[0]
    B = 20000
    counts = plan.sample(probs, B, 3, 0)
    lin = plan.lin(counts, True)
    lin2 = torch.empty_like(lin)
    order = torch.empty((B,), dtype=torch.int32, device="cuda")
    nt.check(lib.qpb_lin_project_ordered(plan.handle, B, nt.ptr(counts), nt.ptr(lin2), nt.ptr(order), nt.stream_ptr()))
    assert torch.equal(lin, lin2)
    o = order.cpu().numpy()
    assert np.array_equal(np.sort(o), np.arange(B))
    raw = plan.lin(counts, False).cpu().numpy()
    raw = raw[..., 0] + 1j * raw[..., 1]
    ev = np.linalg.eigvalsh(raw)
    key = np.where(ev > 0, ev, np.inf).min(axis=1)[o]                     # smallest positive eigenvalue, in start order
    assert np.all(key[1:] >= key[:-1] * 0.8)                              # ascending up to the class width (2^(1/4))
    assert key[0] < 1e-3 < key[-1]
    a, ia = torch.empty_like(lin), torch.empty((B,), dtype=torch.int32, device="cuda")
    b, ib = torch.empty_like(lin), torch.empty((B,), dtype=torch.int32, device="cuda")
    nt.check(lib.qpb_mle_rrr(plan.handle, B, nt.ptr(counts), nt.ptr(lin), 1000, 1e-6, nt.ptr(a), nt.ptr(ia), nt.stream_ptr()))
    nt.check(lib.qpb_mle_rrr_ordered(plan.handle, B, nt.ptr(counts), nt.ptr(lin), nt.ptr(order), 1000, 1e-6, nt.ptr(b),
                                     nt.ptr(ib), nt.stream_ptr()))
    assert torch.equal(a, b) and torch.equal(ia, ib)
    # plans whose kernels take no hint: identity order, same results
    povm3 = qp.generate_measurement_matrix("proj", 3)
    plan3 = engine.state_plan(povm3, np.ones(1) * 10000)
    c3 = plan3.sample(plan3.probabilities(qp.Qobj(haar(3, 1)).bloch)[0], 64, 1, 0)
    l3 = plan3.lin(c3, True)
    l3b, o3 = torch.empty_like(l3), torch.empty((64,), dtype=torch.int32, device="cuda")
    nt.check(lib.qpb_lin_project_ordered(plan3.handle, 64, nt.ptr(c3), nt.ptr(l3b), nt.ptr(o3), nt.stream_ptr()))
    assert torch.equal(l3, l3b) and np.array_equal(o3.cpu().numpy(), np.arange(64))
