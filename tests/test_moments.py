"""MomentInterval and the stats.py moments: oracle against the reference's outputs
(CPU) and the CUDA kernel against both (GPU)."""

import numpy as np
import pytest

from oracle import moments as om

CASES = ("s1", "s2", "s2p")


def test_oracle_matches_reference_moments_and_quantiles(golden):
    g = golden("moments")
    for tag in CASES:
        counts, n = g[f"{tag}_counts"], g[f"{tag}_n_meas"]
        f = counts / n[:, None]
        m, v = om.l2_moments(f, n[0])
        assert abs(m - g[f"{tag}_identity"][0]) < 1e-18 and abs(v - g[f"{tag}_identity"][1]) < 1e-18
        W, dim = om.state_weights(g[f"{tag}_povm"], counts)
        m, v = om.l2_moments(f, n[0], W)
        for distr in ("gamma", "norm", "exp"):
            with np.errstate(invalid="ignore"):
                got = om.quantile(m, v, dim, g["levels"], distr)
            assert np.allclose(got, g[f"{tag}_{distr}"], rtol=0, atol=1e-14, equal_nan=True)
        assert np.allclose(om.quantile(m, v, dim, g["levels"], "gamma", "trace"), g[f"{tag}_gamma_trace"], rtol=0,
                           atol=1e-14)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
def test_kernel_moments_match_reference(golden, tag):
    import quantpy_b200 as qp
    from quantpy_b200 import engine

    g = golden("moments")
    counts, n = g[f"{tag}_counts"], g[f"{tag}_n_meas"]
    f = counts / n[:, None]
    W, dim = om.state_weights(g[f"{tag}_povm"], counts)
    m, v = engine.l2_moments(f, n[0], W)
    mo, vo = om.l2_moments(f, n[0], W)
    assert abs(m - mo) < 1e-12 * abs(mo) and abs(v - vo) < 1e-10 * abs(vo)
    mi, vi = engine.l2_moments(np.stack([f, f]), n[0], om.identity_weights(*f.shape))
    assert np.allclose(mi, g[f"{tag}_identity"][0], rtol=1e-12) and np.allclose(vi, g[f"{tag}_identity"][1], rtol=1e-9)
    n_qubits = int(round(np.log2(g[f"{tag}_povm"].shape[-1]) / 2))
    for dst, key in (("hs", "gamma"), ("trace", "gamma_trace")):
        tmg = qp.StateTomograph(qp.Qobj(np.eye(2**n_qubits) / 2**n_qubits), dst=dst)
        tmg.povm_matrix, tmg.results = g[f"{tag}_povm"], counts
        tmg.n_measurements = n
        dist, cl = qp.MomentInterval(tmg)(g["levels"])
        assert np.abs(dist - g[f"{tag}_{key}"]).max() < 1e-10
    for distr in ("norm", "exp"):
        itv = qp.MomentInterval(tmg.__class__(tmg.state), distr_type=distr)
        itv.tmg.povm_matrix, itv.tmg.results = g[f"{tag}_povm"], counts
        itv.tmg.n_measurements = n
        with np.errstate(invalid="ignore"):
            dist, _ = itv(g["levels"])
        assert np.allclose(dist, g[f"{tag}_{distr}"], rtol=0, atol=1e-10, equal_nan=True)
    with pytest.raises(NotImplementedError):
        qp.MomentInterval(tmg, distr_type="cauchy").setup()


@pytest.mark.gpu
def test_process_moment_interval_matches_reference(golden):
    import quantpy_b200 as qp

    g = golden("moments")
    ptmg = qp.ProcessTomograph(qp.channel.depolarizing(0.1, 1), input_states="sic")
    np.random.seed(0)
    ptmg.experiment(2000, "proj-set")
    ptmg.results = g["p1_counts"]
    dist, _ = qp.MomentInterval(ptmg)(g["levels"])
    assert np.abs(dist - g["p1_gamma"]).max() < 1e-10
    # the analytic radius brackets the bootstrap radius of the same experiment (cross-check named in SURVEY 8f)
    ptmg.point_estimate("lifp", cptp=False)
    boot = qp.BootstrapProcessInterval(ptmg, n_points=4000, cptp=False)
    bdist, _ = boot(np.array([0.5, 0.9]))
    assert np.all(np.abs(bdist / dist[1:3] - 1) < 0.35)
