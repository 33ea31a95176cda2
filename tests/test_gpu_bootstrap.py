"""GPU tests of the fused bootstrap (state) and of the process path."""

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import bootstrap as oboot  # noqa: E402
from oracle import distances as odist  # noqa: E402
from oracle import process as oproc  # noqa: E402
from oracle import state as ostate  # noqa: E402


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


@pytest.mark.parametrize("n,povm,method,dst", [(1, "proj-set", "lin", "hs"), (1, "proj-set", "mle", "trace"),
                                               (2, "proj", "mle", "hs"), (2, "proj", "lin", "if"),
                                               (3, "proj", "lin", "hs")])
def test_fused_bootstrap_is_consistent_with_oracle(qp, n, povm, method, dst):
    """Every stage of the fused call is re-derived by the oracle from the counts the GPU sampled:
    reconstruction and distance agree to 1e-10 sample by sample."""
    from quantpy_b200 import engine

    rho = haar(n, 60 + n)
    tmg = qp.StateTomograph(qp.Qobj(rho), dst=dst)
    np.random.seed(n)
    tmg.experiment(10000, povm)
    plan = engine.state_plan(tmg.povm_matrix, tmg.n_measurements)
    probs = plan.probabilities(tmg.state.bloch)[0]
    B = 400 if n < 3 else 120
    out = plan.bootstrap(probs, B, 11, 0, rho, method=method, max_iter=50, tol=1e-4, dst=dst, keep=True)
    counts = out["counts"].cpu().numpy()
    from quantpy_b200 import _native as nt

    got = nt.complex_to_host(out["rho"])
    if method == "lin":
        want = ostate.lin_estimate(counts, tmg.povm_matrix, tmg.n_measurements)
    else:
        want, wits = ostate.mle_rrr(counts, tmg.povm_matrix, tmg.n_measurements, max_iter=50, tol=1e-4,
                                    return_iters=True)
        assert np.array_equal(out["iters"].cpu().numpy(), wits)
    assert fro(got, want).max() < 1e-10
    # infidelity takes sqrt of eigenvalues: where an estimate is numerically singular (eigenvalue clipped
    # to 1e-15) rounding moves it by ~sqrt(1e-16); see tests/test_oracle_golden.py::sqrtm_tolerances
    tol_d = 1e-10
    if dst == "if":
        tol_d = np.where(np.linalg.eigvalsh(want).min(-1) > 1e-6, 1e-10, 1e-6)
    assert (np.abs(out["dist"].cpu().numpy() - odist.BY_NAME[dst](want, rho)) < tol_d).all()


@pytest.mark.parametrize("method", ["lin", "mle"])
def test_bootstrap_interval_matches_reference_distribution(qp, golden, method):
    """BootstrapStateInterval against the oracle's serial loop (different RNG streams): the quantile
    functions agree within Monte-Carlo error (two-sample KS test)."""
    from scipy import stats

    g = golden("boot_c2_lin")
    tmg = qp.StateTomograph(qp.Qobj(g["rho_true"]))
    tmg.povm_matrix = g["povm_matrix"]
    tmg.results = g["counts"]
    tmg.n_measurements = g["n_meas"]
    tmg.reconstructed_state = qp.Qobj(g["centre"])
    itv = qp.BootstrapStateInterval(tmg, n_points=4000, method=method, tol=1e-6, max_iter=200)
    np.random.seed(3)
    dist, cl = itv()
    assert dist.shape == (1000,) and np.all(np.diff(dist) >= 0)
    assert itv.cl_to_dist.x[0] == 0 and itv.cl_to_dist.x[-1] == 1 and len(itv.cl_to_dist.y) == 4000
    np.random.seed(4)
    ref = oboot.bootstrap_state(g["centre"], g["povm_matrix"], g["n_meas"], 300, method=method, mle="rrr",
                                tol=1e-6, max_iter=200)
    assert stats.ks_2samp(itv.dist, ref).pvalue > 1e-3
    with pytest.raises(ValueError):
        itv.cl_to_dist(1.5)
    # reproducible under np.random.seed, like the reference
    a = qp.BootstrapStateInterval(tmg, n_points=64, method=method)
    b = qp.BootstrapStateInterval(tmg, n_points=64, method=method)
    np.random.seed(8)
    a.setup()
    np.random.seed(8)
    b.setup()
    assert np.array_equal(a.dist, b.dist)


@pytest.mark.parametrize("n,povm,method,dst,B", [(2, "proj", "mle", "hs", 5000), (1, "proj-set", "mle", "hs", 1000),
                                                 (2, "proj-set", "lin", "trace", 3000), (3, "proj", "lin", "if", 600),
                                                 (1, "proj-set", "mle", "hs", 1), (2, "proj", "mle", "hs", 40000)])
def test_one_call_interval_equals_step_by_step_path(qp, n, povm, method, dst, B):
    """qpb_bootstrap_state_interval (host inputs -> quantiles on the host in one library call) against the path it
    replaces (probabilities, fused bootstrap, sort, host interpolation): the same sorted distances, iteration counts
    and quantiles bit for bit; the quantiles also equal the reference expression evaluated in NumPy on `.dist`
    (interval.py:610-612), for the set-up levels and for levels asked later (qpb_quantiles_host)."""
    from scipy.interpolate import interp1d

    from quantpy_b200 import engine

    rho = haar(n, 40 + n)
    tmg = qp.StateTomograph(qp.Qobj(rho), dst=dst)
    np.random.seed(n)
    tmg.experiment(10000, povm)
    tmg.point_estimate("lin")
    levels = np.linspace(1e-3, 1 - 1e-3, 1000)
    res = {}
    for fused in (True, False):
        old = engine.FUSED_INTERVAL
        engine.FUSED_INTERVAL = fused
        try:
            itv = qp.BootstrapStateInterval(tmg, n_points=B, method=method, tol=1e-6, max_iter=300)
            np.random.seed(77)
            q, cl = itv()
            other = itv.cl_to_dist(np.array([0.0, 0.3141, 1.0]))
            res[fused] = (q, other, itv.dist.copy(), itv.iters.copy())
        finally:
            engine.FUSED_INTERVAL = old
    for a, b in zip(res[True], res[False]):
        assert np.array_equal(a, b)
    q, other, dist, iters = res[True]
    assert np.all(np.diff(dist) >= 0) and len(dist) == B
    if B > 1:
        f = interp1d(np.linspace(0, 1, B), dist)
        assert np.abs(q - f(levels)).max() < 1e-15 and np.abs(other - f([0.0, 0.3141, 1.0])).max() < 1e-15
        pos = levels * (B - 1)
        lo = np.minimum(np.floor(pos).astype(np.int64), B - 2)
        assert np.array_equal(q, dist[lo] + (dist[lo + 1] - dist[lo]) * (pos - lo))
    else:
        assert np.array_equal(q, np.full(1000, dist[0]))
    with pytest.raises(ValueError):
        itv.cl_to_dist(-0.1)
    with pytest.raises(ValueError):
        qp.BootstrapStateInterval(tmg, n_points=16, method=method)([0.5, 1.2])


def test_interval_call_without_levels_is_queued_not_lost(qp):
    """qpb_bootstrap_state_interval with n_levels = 0 returns with its work queued (the multi-GPU interval enqueues the
    all-gather behind it).  Calls issued back to back reuse the library's page-locked staging buffer while the
    earlier upload may still be in flight: every call must still see ITS OWN centre state.  Reference: each call
    followed by a synchronisation."""
    import torch

    from quantpy_b200 import engine

    tmg = qp.StateTomograph(qp.Qobj(haar(2, 1)))
    np.random.seed(3)
    tmg.experiment(10000, "proj")
    plan = engine.state_plan(tmg.povm_matrix, tmg.n_measurements)
    centres = [qp.Qobj(haar(2, 50 + i)) for i in range(6)]
    none = np.zeros(0)

    def run(i, sync):
        c = centres[i]
        _, dist, iters = engine.bootstrap_interval(plan, c.bloch, c.matrix, 20000, 11 + i, 0, "mle", True, "lin", 300,
                                                   1e-6, "hs", none)
        if sync:
            torch.cuda.synchronize()
        return dist, iters

    want = [tuple(t.cpu().numpy() for t in run(i, True)) for i in range(len(centres))]
    got = [run(i, False) for i in range(len(centres))]          # six calls in flight on one stream
    q = engine.quantiles_host(got[-1][0], np.array([0.0, 0.5, 1.0]))   # the one synchronisation
    for (d0, i0), (d1, i1) in zip(want, got):
        assert np.array_equal(d0, d1.cpu().numpy()) and np.array_equal(i0, i1.cpu().numpy())
    assert q[0] == want[-1][0][0] and q[2] == want[-1][0][-1]
    assert len({w[0].tobytes() for w in want}) == len(centres)   # the centres really differ


def test_bootstrap_full_size_properties(qp):
    """BASELINE config 2 at full size (1e5 resamples): size-independent properties."""
    from quantpy_b200 import _native as nt
    from quantpy_b200 import engine

    rho = haar(2, 5)
    tmg = qp.StateTomograph(qp.Qobj(rho))
    np.random.seed(0)
    tmg.experiment(10000, "proj")
    plan = engine.state_plan(tmg.povm_matrix, tmg.n_measurements)
    probs = plan.probabilities(tmg.state.bloch)[0]
    B = 100000
    out = plan.bootstrap(probs, B, 1, 0, rho, method="mle", max_iter=200, tol=1e-6, keep=True)
    counts = out["counts"].cpu().numpy()
    assert np.array_equal(counts.sum((1, 2)), np.full(B, 10000))
    est = nt.complex_to_host(out["rho"])
    assert np.abs(np.trace(est, axis1=1, axis2=2) - 1).max() < 1e-12
    assert np.abs(est - est.conj().transpose(0, 2, 1)).max() == 0
    assert np.linalg.eigvalsh(est).min() > -1e-13
    iters = out["iters"].cpu().numpy()
    assert iters.min() >= 1 and iters.max() <= 200
    # the MLE never has a worse likelihood than its linear-inversion start (spot check)
    idx = np.arange(0, B, 997)
    start = ostate.lin_estimate(counts[idx], tmg.povm_matrix, tmg.n_measurements)
    for i, j in enumerate(idx):
        assert ostate.neg_log_likelihood(est[j], counts[j], tmg.povm_matrix, tmg.n_measurements) <= \
            ostate.neg_log_likelihood(start[i], counts[j], tmg.povm_matrix, tmg.n_measurements) + 1e-12
    # a prefix of the batch reproduces exactly (global-index keyed RNG, deterministic kernels)
    again = plan.bootstrap(probs, 1000, 1, 0, rho, method="mle", max_iter=200, tol=1e-6)
    assert np.array_equal(again["dist"].cpu().numpy(), out["dist"].cpu().numpy()[:1000])
    d = out["dist"].cpu().numpy()
    assert np.abs(d[idx] - odist.hs(est[idx], rho)).max() < 1e-12


def test_bootstrap_custom_distance_and_errors(qp):
    rho = haar(1, 9)
    tmg = qp.StateTomograph(qp.Qobj(rho), dst=lambda a, b: float(np.abs(a.matrix - b.matrix).max()))
    np.random.seed(0)
    tmg.experiment(1000, "proj-set")
    itv = qp.BootstrapStateInterval(tmg, n_points=50)
    itv.setup()
    assert itv.dist.shape == (50,) and itv.dist.max() < 0.2
    ch = qp.ProcessTomograph(qp.channel.depolarizing(0.1, 1), "sic")
    with pytest.raises(NotImplementedError):
        qp.BootstrapStateInterval(ch).setup()
    with pytest.raises(NotImplementedError):
        qp.BootstrapProcessInterval(tmg).setup()


# ----------------------------------------------------------------------------- process path

def process_tomograph(qp, g, n):
    chan = qp.channel.depolarizing(0.1, n)
    tmg = qp.ProcessTomograph(chan, "sic")
    np.random.seed(0)
    tmg.experiment(10000, "proj-set")
    return tmg


def test_process_input_json(qp, golden):
    """The reference's only shipped known-input run (input.json + scripts/process_interval.py --no-ci)."""
    g = golden("process")
    inputs = [qp.Qobj(m) for m in g["json_inputs"]]
    tmg = qp.ProcessTomograph(qp.channel.depolarizing(n_qubits=1), input_states=inputs)
    np.random.seed(0)
    tmg.experiment(1000, "proj-set")
    tmg.results = g["json_outcomes"]
    est = tmg.point_estimate(cptp=False)
    assert fro(est.choi.matrix, g["json_choi"]) < 1e-10
    assert np.abs(est.choi.bloch - g["json_choi_bloch"]).max() < 1e-10
    assert fro(tmg.point_estimate(cptp=True).choi.matrix, g["json_choi_cptp"]) < 1e-9


@pytest.mark.parametrize("n", [1, 2])
def test_process_lifp_and_cptp_match_reference(qp, golden, n):
    g = golden("process")
    tmg = process_tomograph(qp, g, n)
    counts = g[f"dep{n}_counts"]
    assert tmg.results.shape == counts.shape[1:]
    raw = tmg.point_estimate_batch(counts, cptp=False)
    assert fro(raw, g[f"dep{n}_lifp_raw"]).max() < 1e-10
    proj, iters = tmg.point_estimate_batch(counts, cptp=True, return_iters=True)
    want, wit = oproc.cptp_projection(g[f"dep{n}_lifp_raw"], return_iters=True)
    assert fro(proj, g[f"dep{n}_lifp_cptp"]).max() < 1e-9
    assert fro(proj, want).max() < 1e-9 and np.abs(iters - wit).max() <= 1
    tmg.results = counts[0]
    st = tmg.point_estimate("states", cptp=True)
    assert fro(st.choi.matrix, g[f"dep{n}_states_cptp"][0]) < 1e-9
    one = tmg.cptp_projection(qp.Channel(g[f"dep{n}_lifp_raw"][0]))
    assert fro(one.choi.matrix, want[0]) < 1e-9
    with pytest.raises(ValueError):
        tmg.point_estimate("nope")


def test_process_bootstrap(qp, golden):
    from scipy import stats

    g = golden("process")
    tmg = process_tomograph(qp, g, 1)
    tmg.results = g["boot1_counts"]
    tmg.reconstructed_channel = qp.Channel(g["boot1_centre"])
    itv = qp.BootstrapProcessInterval(tmg, n_points=2000)
    np.random.seed(1)
    dist, cl = itv([0.1, 0.5, 0.9])
    assert np.all(np.diff(dist) > 0)
    inputs = oproc.input_states("sic", 1)
    np.random.seed(2)
    ref = oboot.bootstrap_process(g["boot1_centre"], inputs, g["dep1_povm"], g["dep1_n_meas"], 200)
    assert stats.ks_2samp(itv.dist, ref).pvalue > 1e-3
    # counts sampled for the bootstrap are consistent with the channel outputs
    counts = tmg.sample_counts(300, 10000, "proj-set", seed=3)
    assert counts.shape == (300, 4, 3, 2) and (counts.sum(-1) == 10000).all()
    probs = np.array([ostate.probabilities(g["dep1_povm"], __import__("oracle").pauli.matrix_to_bloch(o))
                      for o in g["dep1_outputs"]])
    assert np.abs(counts.mean(0) / 10000 - probs).max() < 5e-3


# ----------------------------------------------------------------------------- BASELINE configs at full size

def _plan_for(qp, n, povm, seed):
    from quantpy_b200 import engine

    rho = haar(n, seed)
    pm = qp.generate_measurement_matrix(povm, n)
    plan = engine.state_plan(pm, np.ones(pm.shape[0]) * 10000)
    return rho, pm, plan, plan.probabilities(qp.Qobj(rho).bloch)[0]


def test_config1_one_qubit_bootstrap_full_size(qp):
    """BASELINE configs[0]: 1 qubit, 'proj-set', 10k shots, MLE estimate + 1000-sample bootstrap CI."""
    rho = haar(1, 11)
    tmg = qp.StateTomograph(qp.Qobj(rho))
    np.random.seed(0)
    tmg.experiment(10000)
    est = tmg.point_estimate("mle")
    assert qp.hs_dst(est, qp.Qobj(rho)) < 0.02
    itv = qp.BootstrapStateInterval(tmg, n_points=1000, method="mle")
    dist, cl = itv()
    assert dist.shape == cl.shape == (1000,) and np.all(np.diff(dist) >= 0)
    assert itv.state is tmg.reconstructed_state  # centre = the reconstructed state, interval.py:587
    # the true state lies within the 99.9 % radius around the estimate
    assert qp.hs_dst(est, qp.Qobj(rho)) < itv.cl_to_dist(0.999) * 1.5
    assert 0.002 < itv.cl_to_dist(0.5) < 0.02


def test_config3_three_qubit_lin_full_size(qp):
    """BASELINE configs[2]: 3 qubits, 'lin' + projection (Jacobi), 1e5 resamples: physicality, idempotence,
    and agreement with the oracle on a strided subset."""
    from quantpy_b200 import _native as nt

    rho, pm, plan, probs = _plan_for(qp, 3, "proj", 3)
    B = 100000
    out = plan.bootstrap(probs, B, 5, 0, rho, method="lin", dst="hs", keep=True)
    est = nt.complex_to_host(out["rho"])
    counts = out["counts"].cpu().numpy()
    assert np.array_equal(counts.sum((1, 2)), np.full(B, 10000))
    assert np.abs(np.trace(est, axis1=1, axis2=2) - 1).max() < 1e-12
    assert np.abs(est - est.conj().transpose(0, 2, 1)).max() < 1e-15
    assert np.linalg.eigvalsh(est).min() >= 0.5e-15  # clipped at 1e-15, then renormalised
    idx = np.arange(0, B, 1009)
    want = ostate.lin_estimate(counts[idx], pm, np.ones(1) * 10000)
    assert fro(est[idx], want).max() < 1e-10
    assert np.abs(out["dist"].cpu().numpy()[idx] - odist.hs(want, rho)).max() < 1e-10
    # projecting an already physical state changes nothing (idempotence of _make_feasible)
    again = ostate.make_feasible(est[idx])
    assert fro(again, est[idx]).max() < 1e-12


def test_config4_four_qubit_mle_full_size(qp):
    """BASELINE configs[3]: 4 qubits (d = 16, 1296 outcomes), MLE, 1e4 resamples."""
    from quantpy_b200 import _native as nt

    rho, pm, plan, probs = _plan_for(qp, 4, "proj", 4)
    B = 10000
    out = plan.bootstrap(probs, B, 6, 0, rho, method="mle", max_iter=20, tol=1e-6, dst="hs", keep=True)
    est = nt.complex_to_host(out["rho"])
    counts = out["counts"].cpu().numpy()
    assert np.array_equal(counts.sum((1, 2)), np.full(B, 10000))
    assert np.abs(np.trace(est, axis1=1, axis2=2) - 1).max() < 1e-12
    assert np.linalg.eigvalsh(est).min() > -1e-13
    idx = np.array([0, 4999, 9999])
    want, wits = ostate.mle_rrr(counts[idx], pm, np.ones(1) * 10000, max_iter=20, tol=1e-6, return_iters=True)
    assert np.array_equal(out["iters"].cpu().numpy()[idx], wits)
    assert fro(est[idx], want).max() < 1e-10
    for i, j in enumerate(idx):  # likelihood never decreases from the linear-inversion start
        start = ostate.lin_estimate(counts[j], pm, np.ones(1) * 10000)
        assert ostate.neg_log_likelihood(est[j], counts[j], pm, np.ones(1) * 10000) <= \
            ostate.neg_log_likelihood(start, counts[j], pm, np.ones(1) * 10000) + 1e-12


@pytest.mark.parametrize("n,B", [(1, 20000), (2, 600)])
def test_config5_process_bootstrap_full_size(qp, n, B):
    """BASELINE configs[4]: depolarising channel, lifp + CPTP projection: every bootstrap estimate is CPTP,
    and a strided subset matches the oracle."""
    chan = qp.channel.depolarizing(0.1, n)
    tmg = qp.ProcessTomograph(chan, "sic")
    np.random.seed(0)
    tmg.experiment(10000, "proj-set")
    counts = tmg.sample_counts(B, 10000, "proj-set", seed=9)
    choi, iters = tmg.point_estimate_batch(counts, cptp=True, return_iters=True)
    d = 2**n
    rho_in = np.einsum("niaja->nij", choi.reshape(B, d, d, d, d))
    assert np.abs(rho_in - np.eye(d)).max() < 1e-5       # trace preserving (Channel.is_cptp's tolerance)
    assert np.linalg.eigvalsh(0.5 * (choi + choi.conj().transpose(0, 2, 1))).min() > -1e-5  # completely positive
    idx = np.arange(0, B, max(1, B // 7))
    inputs = oproc.input_states("sic", n)
    pm = qp.generate_measurement_matrix("proj-set", n)
    want, wit = oproc.lifp_estimate(counts[idx], inputs, pm, np.ones(pm.shape[0]) * 10000, cptp=True, return_iters=True)
    assert fro(choi[idx], want).max() < 1e-9 and np.abs(iters[idx] - wit).max() <= 1
    itv = qp.BootstrapProcessInterval(tmg, n_points=B // 2, channel=chan)
    itv.setup(seed=3)
    assert np.all(np.diff(itv.dist) >= 0) and itv.dist[0] > 0


@pytest.mark.parametrize("n,method", [(1, "lin"), (1, "mle"), (2, "lin")])
def test_process_states_method_batched(qp, golden, n, method):
    """'states' estimates with per-state reconstructions batched on the state kernels, Choi
    assembly and conditional CPTP projection on the device, against the oracle's restatement of process.py:316-327."""
    g = golden("process")
    tmg = process_tomograph(qp, g, n)
    counts = g[f"dep{n}_counts"]
    if method == "lin":
        tmg.results = counts[0]
        one = tmg.point_estimate("states", cptp=True)
        assert fro(one.choi.matrix, g[f"dep{n}_states_cptp"][0]) < 1e-9  # reference output
    inputs = oproc.input_states("sic", n)
    pm, n_meas = g[f"dep{n}_povm"], g[f"dep{n}_n_meas"]
    more = tmg.sample_counts(6, 10000, "proj-set", seed=4)
    kw = dict(max_iter=30, tol=1e-6)
    for cptp in (False, True):
        got, iters = tmg.point_estimate_states_batch(more, cptp=cptp, method=method, n_iter=30, tol=1e-6,
                                                     return_iters=True)
        for b in range(len(more)):
            want = oproc.states_estimate(more[b], inputs, pm, n_meas, cptp=cptp, method=method, mle="rrr", **kw)
            assert fro(got[b], want) < 1e-9
            assert (iters[b] == 0) == (not cptp or oproc.is_cptp(oproc.states_estimate(
                more[b], inputs, pm, n_meas, cptp=False, method=method, mle="rrr", **kw)))
    itv = qp.BootstrapProcessInterval(tmg, n_points=300, method="states", states_est_method=method,
                                      channel=tmg.channel)
    itv.setup(seed=1)
    assert itv.dist.shape == (300,) and np.all(np.diff(itv.dist) >= 0) and itv.dist[0] > 0


def test_start_order_and_hand_over_policies_keep_every_bit(qp):
    """The fused bootstrap starts the likely long runners first (order from the linear estimate's smallest eigenvalue)
    and moves long-running samples between lane mappings; none of that may change a bit of any distance or
    iteration count: every policy equals the plain thread-per-sample launch in index order."""
    from quantpy_b200 import _native as nt
    from quantpy_b200 import engine

    rho = haar(2, 3)
    povm = qp.generate_measurement_matrix("proj", 2)
    plan = engine.state_plan(povm, np.ones(1) * 10000)
    probs = plan.probabilities(qp.Qobj(rho).bloch)[0].contiguous()
    ref = nt.complex_to_device(rho)
    for B in (5000, 60000):
        bufs = plan.bootstrap_buffers(B)
        with nt.option("NO_TAIL_MERGE", 1), nt.option("NO_MLE_ORDER", 1):
            plan.bootstrap_into(bufs, probs, ref, 11, 0, method="mle", max_iter=1000, tol=1e-6)
        want_d, want_it = bufs["dist"].clone(), bufs["iters"].clone()
        settings = [{}, {"NO_MLE_ORDER": 1}, {"MLE_MERGE": 1}, {"MLE_TAIL_POLL": 4, "MLE_ADOPT": 8, "MLE_PARK_LIVE": 12},
                    {"MLE_PARK_AGE": 150, "MLE_PARK_AGE_LO": 60, "MLE_PARK_AGE_END": 80},
                    {"MLE_MERGE": 4, "MLE_TAIL_POLL": 2, "MLE_TAIL_AGE": 32, "MLE_ADOPT": 1}]
        for opts in settings:
            import contextlib
            with contextlib.ExitStack() as stack:
                for k, v in opts.items():
                    stack.enter_context(nt.option(k, v))
                bufs["dist"].zero_()
                bufs["iters"].zero_()
                plan.bootstrap_into(bufs, probs, ref, 11, 0, method="mle", max_iter=1000, tol=1e-6)
            assert bool((bufs["dist"] == want_d).all()) and bool((bufs["iters"] == want_it).all()), (B, opts)
        assert int(want_it.max()) > 300


def test_pauli2_scheduling_stress(qp):
    """tools/stress_pauli2.py in short: 80 fused launches back to back with random batch sizes (around every policy
    threshold), state ranks, tolerances, iteration caps and scheduling options; each equals the plain
    thread-per-sample launch of the same inputs bit for bit (and none hangs: pytest-timeout is the guard)."""
    import subprocess
    import sys as _sys

    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "stress_pauli2.py")
    out = subprocess.run([_sys.executable, tool, "7", "80"], capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "80 launches, 0 mismatches" in out.stdout, out.stdout[-2000:]


@pytest.mark.parametrize("n", [1, 2, 3, 31, 1000, 2047, 2048, 2049, 4097, 12500, 16384, 16385, 25000, 50000, 100000,
                               131072, 131073, 333333, 1048576, 1048577])
def test_sort_kernels_match_numpy(qp, n):
    """qpb_sort_f64: one CTA's bitonic network up to 2048 keys, four-launch sample sort up to 1M keys, device
    radix sort above; keys only, bit-exact."""
    import torch

    from quantpy_b200 import engine

    rng = np.random.default_rng(n)
    x = rng.random(n)
    x[: n // 7] = x[n // 5: n // 5 + n // 7]  # ties
    got = engine.sort_f64(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(got, np.sort(x))


@pytest.mark.parametrize("kind", ["constant", "two-values", "heavy-zero", "sorted", "reversed", "signed", "skewed",
                                  "one-big-gap", "one-huge-gap"])
def test_sample_sort_on_adversarial_inputs(qp, kind):
    """The sample sort's buckets come from sampled splitters: duplicates (their own 'equal' buckets), monotone
    inputs, negative keys, heavy skew and an oversized bucket (the counting fallback) all give np.sort's result, and
    the same array as the radix sort it replaces."""
    import torch

    from quantpy_b200 import _native as nt
    from quantpy_b200 import engine

    n = 100000
    rng = np.random.default_rng(1)
    # the splitter kernel reads 4 * 256 keys of a 1e5-key input at these positions; keys placed elsewhere are
    # invisible to it
    seen = np.floor((np.arange(1024) + 0.5) * n / 1024).astype(np.int64)
    n_hidden = 6000 if kind == "one-huge-gap" else 3000  # ONE open bucket (0, 1): counting fallback / bitonic network
    hidden = rng.choice(np.setdiff1d(np.arange(n), seen), n_hidden, replace=False)
    gap = rng.integers(0, 2, n).astype(float)
    gap[hidden] = 0.5 + 1e-9 * rng.random(n_hidden)
    x = {
        "constant": np.full(n, 0.25),
        "two-values": rng.integers(0, 2, n).astype(float),
        "heavy-zero": np.where(rng.random(n) < 0.6, 0.0, rng.random(n)),
        "sorted": np.sort(rng.random(n)),
        "reversed": np.sort(rng.random(n))[::-1].copy(),
        "signed": rng.normal(size=n) * 10.0 ** rng.integers(-200, 200, n),
        "skewed": rng.random(n) ** 40,
        "one-big-gap": gap,
        "one-huge-gap": gap,
    }[kind]
    dev = torch.from_numpy(x).cuda()
    got = engine.sort_f64(dev).cpu().numpy()
    assert np.array_equal(got, np.sort(x))
    with nt.option("NO_SAMPLE_SORT", 1):
        assert np.array_equal(engine.sort_f64(dev).cpu().numpy(), got)


@pytest.mark.parametrize("lens", [[5], [4, 3], [12500] * 8, [100000, 99999, 1, 0, 7], [0, 0, 3]])
def test_merge_sorted_runs_is_the_sorted_concatenation(qp, lens):
    """qpb_merge_sorted_runs (the multi-GPU quantile step): padded, individually sorted shards -> the bit-exact
    `dist.sort()` of their concatenation (interval.py:610), including ties across runs."""
    import ctypes

    import torch

    from quantpy_b200 import _native as nt

    rng = np.random.default_rng(len(lens))
    width = max(max(lens), 1)
    runs = [np.sort(np.round(rng.random(m), 3)) for m in lens]  # rounding makes cross-run ties common
    buf = np.full(len(lens) * width, -1.0)
    for r, run in enumerate(runs):
        buf[r * width: r * width + len(run)] = run
    total = int(sum(lens))
    dev = torch.from_numpy(buf).cuda()
    out = torch.full((total,), np.nan, dtype=torch.float64, device="cuda")
    ln = np.asarray(lens, dtype=np.int32)
    st = np.arange(len(lens), dtype=np.int64) * width
    nt.check(nt.load_library().qpb_merge_sorted_runs(len(lens), ln.ctypes.data_as(ctypes.c_void_p),
                                                     st.ctypes.data_as(ctypes.c_void_p), nt.ptr(dev), nt.ptr(out),
                                                     nt.stream_ptr()))
    assert np.array_equal(out.cpu().numpy(), np.sort(np.concatenate(runs)))


def test_tma_pipelined_gemm_equals_plain_gemm(qp):
    """k_gemm_counts_dmma (TMA bulk copies into a 3-stage mbarrier ring, persistent CTAs) against k_gemm_counts_simple:
    same DMMA accumulation order, so linear inversion (n = 3, 4: M, K tails) and 'lifp' (grouped normalisation,
    N = 32 and 512) must agree in every bit; both are separately held to 1e-10 against the oracle elsewhere."""
    import torch

    from quantpy_b200 import _native as nt
    from quantpy_b200 import engine

    for n, B in ((3, 1000 + 37), (4, 130)):
        rho, pm, plan, probs = _plan_for(qp, n, "proj", 20 + n)
        counts = plan.sample(probs, B, 9, 0)
        a = plan.lin(counts, False).cpu().numpy()
        with nt.option("NO_TMA_GEMM", 1):
            b = plan.lin(counts, False).cpu().numpy()
        assert np.array_equal(a, b), n
    for n, B in ((1, 333), (2, 70)):
        chan = qp.channel.depolarizing(0.1, n)
        tmg = qp.ProcessTomograph(chan, "sic")
        np.random.seed(n)
        tmg.experiment(10000, "proj-set")
        counts = tmg.sample_counts(B, 10000, "proj-set", seed=4, device=True)
        a = tmg.point_estimate_batch(counts, cptp=False)
        with nt.option("NO_TMA_GEMM", 1):
            b = tmg.point_estimate_batch(counts, cptp=False)
        assert np.array_equal(a, b), n
