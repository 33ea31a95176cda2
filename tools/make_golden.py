#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (nordmtr/quantpy).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tools/make_golden.py

The reference imports ``cvxopt`` at module scope (quantpy/tomography/interval.py:6) and
cvxopt is not installed here, so a stub module is registered first; nothing on the
state/process bootstrap path touches it.  Every array stored below is an OUTPUT OF THE
REFERENCE for the stored inputs; the tests compare the oracle and the CUDA path to them.
"""

import json
import os
import sys
import types
import warnings

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference():
    if "cvxopt" not in sys.modules:
        stub = types.ModuleType("cvxopt")

        class _Solvers:
            options = {}

        stub.matrix = lambda *a, **k: None
        stub.solvers = _Solvers()
        sys.modules["cvxopt"] = stub
    sys.path.insert(0, REF)
    import quantpy  # noqa: E402

    return quantpy


def haar_mixed(n, rng, rank=None):
    d = 2**n
    k = d if rank is None else rank
    g = rng.normal(size=(d, k)) + 1j * rng.normal(size=(d, k))
    rho = g @ g.conj().T
    return rho / np.trace(rho)


def state_case(qp, n, povm, n_shots, seed, n_samples, rank=None, with_mle=True, mle_tight=False):
    rng = np.random.default_rng(seed)
    rho = haar_mixed(n, rng, rank)
    tmg = qp.StateTomograph(qp.Qobj(rho))
    np.random.seed(seed)
    out = {"rho_true": rho, "n_qubits": n, "povm_name": povm, "seed": seed}
    counts, lin_phys, lin_raw, mle_def, mle_tightv, n_meas = [], [], [], [], [], None
    d_hs, d_tr, d_if = [], [], []
    for _ in range(n_samples):
        tmg.experiment(n_shots, povm)
        counts.append(tmg.results.copy())
        n_meas = np.asarray(tmg.n_measurements, dtype=float)
        lp = tmg.point_estimate("lin", physical=True).matrix
        lr = tmg.point_estimate("lin", physical=False).matrix
        lin_phys.append(lp)
        lin_raw.append(lr)
        d_hs.append(float(qp.hs_dst(lp, rho)))
        d_tr.append(float(qp.trace_dst(lp, rho)))
        d_if.append(float(qp.if_dst(lp, rho)))
        if with_mle:
            mle_def.append(tmg.point_estimate("mle").matrix)
        if mle_tight:
            mle_tightv.append(tmg.point_estimate("mle", tol=1e-12, max_iter=100000).matrix)
    out.update(
        povm_matrix=tmg.povm_matrix, n_meas=n_meas, counts=np.array(counts),
        lin_physical=np.array(lin_phys), lin_raw=np.array(lin_raw),
        dist_hs=np.array(d_hs), dist_trace=np.array(d_tr), dist_if=np.array(d_if),
        probs=np.clip(np.einsum("ijk,k->ij", tmg.povm_matrix, tmg.state.bloch) * 2**n, 0, 1),
    )
    if with_mle:
        out["mle_default"] = np.array(mle_def)
    if mle_tight:
        out["mle_tight"] = np.array(mle_tightv)
    return out


def bootstrap_case(qp, n, povm, n_shots, seed, n_points, method):
    rng = np.random.default_rng(seed)
    rho = haar_mixed(n, rng)
    tmg = qp.StateTomograph(qp.Qobj(rho))
    np.random.seed(seed)
    tmg.experiment(n_shots, povm)
    first_counts = tmg.results.copy()
    centre = tmg.point_estimate("lin").matrix
    itv = qp.BootstrapStateInterval(tmg, n_points=n_points, method=method)
    np.random.seed(seed + 1)
    dist, cl = itv(np.array([0.0, 0.1, 0.5, 0.9, 0.95, 1.0]))
    xs = itv.cl_to_dist.x
    ys = itv.cl_to_dist.y
    return dict(rho_true=rho, counts=first_counts, centre=centre, povm_matrix=tmg.povm_matrix,
                n_meas=np.asarray(tmg.n_measurements, float), sorted_dist=ys, grid=xs,
                query_cl=cl, query_dist=dist, seed=seed, n_points=n_points)


def process_cases(qp):
    out = {}
    # --- the only known-input run the reference ships: input.json + scripts/process_interval.py --no-ci
    with open(os.path.join(REF, "input.json")) as fp:
        data = json.load(fp)
    results = np.asarray(data["outcomes"])
    povm = np.asarray(data["povm_matrix"])
    inputs = [qp.Qobj(b) for b in data["input_states"]]
    tmg = qp.ProcessTomograph(qp.channel.depolarizing(n_qubits=1), input_states=inputs)
    np.random.seed(0)
    tmg.experiment(1000, "proj-set")
    tmg.results = results
    ch = tmg.point_estimate(cptp=False)
    out["json_outcomes"] = results
    out["json_povm"] = povm
    out["json_inputs"] = np.array([q.matrix for q in inputs])
    out["json_choi_bloch"] = np.asarray(ch.choi.bloch)
    out["json_choi"] = np.asarray(ch.choi.matrix)
    out["json_choi_cptp"] = np.asarray(tmg.point_estimate(cptp=True).choi.matrix)
    # --- depolarising channel, 1 and 2 qubits, sic inputs
    for n, nrep in ((1, 4), (2, 2)):
        chan = qp.channel.depolarizing(p=0.1, n_qubits=n)
        tmg = qp.ProcessTomograph(chan, input_states="sic")
        np.random.seed(10 + n)
        counts, raw, proj, sts, d_hs = [], [], [], [], []
        for _ in range(nrep):
            tmg.experiment(10000, "proj-set")
            counts.append(tmg.results.copy())
            raw.append(tmg.point_estimate("lifp", cptp=False).choi.matrix)
            est = tmg.point_estimate("lifp", cptp=True)
            proj.append(est.choi.matrix)
            d_hs.append(float(qp.hs_dst(est.choi, chan.choi)))
            sts.append(tmg.point_estimate("states", cptp=True).choi.matrix)
        out[f"dep{n}_choi_true"] = np.asarray(chan.choi.matrix)
        out[f"dep{n}_inputs"] = np.array([q.matrix for q in tmg.input_basis.elements])
        out[f"dep{n}_outputs"] = np.array([t.state.matrix for t in tmg.tomographs])
        out[f"dep{n}_povm"] = tmg.tomographs[0].povm_matrix
        out[f"dep{n}_n_meas"] = np.asarray(tmg.tomographs[0].n_measurements, float)
        out[f"dep{n}_counts"] = np.array(counts)
        out[f"dep{n}_lifp_raw"] = np.array(raw)
        out[f"dep{n}_lifp_cptp"] = np.array(proj)
        out[f"dep{n}_states_cptp"] = np.array(sts)
        out[f"dep{n}_dist_hs"] = np.array(d_hs)
    # --- 1-qubit process bootstrap with the legacy RNG stream
    chan = qp.channel.depolarizing(p=0.1, n_qubits=1)
    tmg = qp.ProcessTomograph(chan, input_states="sic")
    np.random.seed(21)
    tmg.experiment(10000, "proj-set")
    tmg.point_estimate("lifp", cptp=True)
    itv = qp.BootstrapProcessInterval(tmg, n_points=12)
    np.random.seed(22)
    itv.setup()
    out["boot1_counts"] = tmg.results.copy()
    out["boot1_centre"] = np.asarray(tmg.reconstructed_channel.choi.matrix)
    out["boot1_sorted_dist"] = itv.cl_to_dist.y
    return out


def api_cases(qp):
    out = {}
    for name in ("proj", "proj-set", "proj4", "sic"):
        for n in (1, 2):
            out[f"povm_{name}_{n}"] = qp.generate_measurement_matrix(name, n)
    rng = np.random.default_rng(5)
    for n in (1, 2, 3):
        rho = haar_mixed(n, rng)
        out[f"rho_{n}"] = rho
        out[f"bloch_{n}"] = qp.Qobj(rho).bloch
        out[f"back_{n}"] = qp.Qobj(qp.Qobj(rho).bloch).matrix
    rho2 = out["rho_2"]
    out["ptrace_keep0"] = qp.Qobj(rho2).ptrace([0]).matrix
    out["ptrace_keep1"] = qp.Qobj(rho2).ptrace([1]).matrix
    chan = qp.channel.depolarizing(p=0.3, n_qubits=1)
    out["dep03_choi"] = chan.choi.matrix
    out["dep03_apply"] = qp.Channel(chan.choi.matrix).transform(qp.Qobj(out["rho_1"])).matrix
    out["zchan_choi"] = qp.operator.Z.as_channel().choi.matrix
    out["ket01_bloch"] = qp.Qobj([0, 1], is_ket=True).bloch
    basis = qp.basis.Basis([qp.Qobj(b) for b in np.squeeze(qp.generate_measurement_matrix("sic", 1))])
    out["basis_gram"] = basis.gram
    out["basis_decomp"] = basis.decompose(qp.Qobj(out["rho_1"]))
    a, b = haar_mixed(2, rng), haar_mixed(2, rng)
    out["dst_a"], out["dst_b"] = a, b
    out["dst_vals"] = np.array([qp.hs_dst(a, b), qp.trace_dst(a, b), qp.if_dst(a, b)], dtype=float)
    return out


def main():
    warnings.filterwarnings("ignore")
    os.makedirs(OUT, exist_ok=True)
    qp = load_reference()
    cases = {
        "state_c1": state_case(qp, 1, "proj-set", 10000, 1, 6, mle_tight=True),
        "state_c1_pure": state_case(qp, 1, "proj-set", 10000, 2, 4, rank=1, mle_tight=True),
        "state_c2": state_case(qp, 2, "proj", 10000, 3, 6, mle_tight=True),
        "state_c2_set": state_case(qp, 2, "proj-set", 10000, 4, 4),
        "state_c2_rank1": state_case(qp, 2, "proj", 10000, 5, 4, rank=1),
        "state_c2_sic": state_case(qp, 2, "sic", 10000, 6, 4),
        "state_c3": state_case(qp, 3, "proj", 10000, 7, 3, with_mle=False),
        "state_c3_rank2": state_case(qp, 3, "proj", 10000, 8, 3, rank=2, with_mle=False),
        "state_c4": state_case(qp, 4, "proj", 10000, 9, 2, with_mle=False),
        "boot_c1_lin": bootstrap_case(qp, 1, "proj-set", 10000, 31, 40, "lin"),
        "boot_c2_lin": bootstrap_case(qp, 2, "proj", 10000, 32, 25, "lin"),
        "boot_c1_mle": bootstrap_case(qp, 1, "proj-set", 10000, 33, 10, "mle"),
        "process": process_cases(qp),
        "api": api_cases(qp),
        "polytopes": polytope_cases(qp),
        "moments": moment_cases(qp),
    }
    for name, arrays in cases.items():
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(f"{name}: {os.path.getsize(path)} bytes")


if __name__ == "__main__" and not {"--polytopes", "--moments", "--mle-constr", "--mhmc"} & set(sys.argv):
    main()


def polytope_cases(qp):
    """quantpy/tomography/polytopes: deterministic pieces on reference-sampled counts, plus two seeded
    coverage runs (Monte-Carlo figures, compared statistically)."""
    from quantpy.tomography.polytopes import utils as putils
    from quantpy.tomography.polytopes import verification as pver

    out = {}
    levels = np.array([0.0, 0.5, 0.9, 0.99])
    out["levels"] = levels
    for n, tag in ((1, "q1"), (2, "q2")):
        rng = np.random.default_rng(40 + n)
        rho = haar_mixed(n, rng)
        tmg = qp.StateTomograph(qp.Qobj(rho))
        np.random.seed(50 + n)
        counts, deltas, confs = [], [], []
        for _ in range(4):
            tmg.experiment(1000)
            counts.append(tmg.results.copy())
            f = np.clip(tmg.results / tmg.n_measurements[:, None], 1e-15, 1 - 1e-15)
            deltas.append([putils.count_delta(cl, f, tmg.n_measurements) for cl in levels])
            confs.append([putils.count_confidence(d, f, tmg.n_measurements) for d in (1e-3, 0.02, 0.1)])
        out[f"{tag}_rho"] = rho
        out[f"{tag}_povm"] = tmg.povm_matrix
        out[f"{tag}_n_meas"] = np.asarray(tmg.n_measurements, float)
        out[f"{tag}_counts"] = np.array(counts)
        out[f"{tag}_deltas"] = np.array(deltas)
        out[f"{tag}_confs"] = np.array(confs)
    # seeded coverage experiments (tqdm progress goes to stderr)
    np.random.seed(60)
    out["qst_state"] = haar_mixed(1, np.random.default_rng(61))
    out["qst_cover"] = pver.test_qst(qp.Qobj(out["qst_state"]), levels[1:], n_measurements=1000, n_trials=300)
    np.random.seed(62)
    out["qpt_cover"] = pver.test_qpt(qp.channel.depolarizing(0.1, 1), levels[1:], n_measurements=1000, n_trials=100)
    return out


if __name__ == "__main__" and "--polytopes" in sys.argv:
    warnings.filterwarnings("ignore")
    arrays = polytope_cases(load_reference())
    path = os.path.join(OUT, "polytopes.npz")
    np.savez_compressed(path, **arrays)
    print(f"polytopes: {os.path.getsize(path)} bytes")


def moment_cases(qp):
    """quantpy/stats.py moments and MomentInterval quantiles on reference-sampled counts."""
    from quantpy import stats as qstats

    out = {"levels": np.array([0.1, 0.5, 0.9, 0.99])}
    for n, povm, tag in ((1, "proj-set", "s1"), (2, "proj-set", "s2"), (2, "proj", "s2p")):
        rho = haar_mixed(n, np.random.default_rng(70 + n))
        tmg = qp.StateTomograph(qp.Qobj(rho))
        np.random.seed(71 + n)
        tmg.experiment(2000, povm)
        freq = tmg.results / tmg.n_measurements[:, None]
        out[f"{tag}_povm"], out[f"{tag}_counts"], out[f"{tag}_n_meas"] = tmg.povm_matrix, tmg.results, np.asarray(tmg.n_measurements, float)
        out[f"{tag}_identity"] = np.array([qstats.l2_mean(freq, 2000), qstats.l2_variance(freq, 2000)])
        for distr in ("gamma", "norm", "exp"):
            itv = qp.MomentInterval(tmg, distr_type=distr)
            itv.setup()
            out[f"{tag}_{distr}"] = itv.cl_to_dist(out["levels"])
        tmg_tr = qp.StateTomograph(qp.Qobj(rho), dst="trace")
        tmg_tr.povm_matrix, tmg_tr.results = tmg.povm_matrix, tmg.results
        tmg_tr.n_measurements = tmg.n_measurements
        itv = qp.MomentInterval(tmg_tr)
        itv.setup()
        out[f"{tag}_gamma_trace"] = itv.cl_to_dist(out["levels"])
    chan = qp.channel.depolarizing(0.1, 1)
    ptmg = qp.ProcessTomograph(chan, input_states="sic")
    np.random.seed(75)
    ptmg.experiment(2000, "proj-set")
    itv = qp.MomentInterval(ptmg)
    itv.setup()
    out["p1_counts"] = ptmg.results
    out["p1_gamma"] = itv.cl_to_dist(out["levels"])
    return out


if __name__ == "__main__" and "--moments" in sys.argv:
    warnings.filterwarnings("ignore")
    arrays = moment_cases(load_reference())
    path = os.path.join(OUT, "moments.npz")
    np.savez_compressed(path, **arrays)
    print(f"moments: {os.path.getsize(path)} bytes")


def mle_constr_cases(qp):
    """Reference 'mle-constr' (SLSQP with the unit-trace equality, state.py:231-254) on the counts already stored
    in state_c1 / state_c2_set, at the default and at a tight tolerance."""
    out = {}
    for tag in ("state_c1", "state_c2_set"):
        g = np.load(os.path.join(OUT, tag + ".npz"))
        tmg = qp.StateTomograph(qp.Qobj(g["rho_true"]))
        tmg.povm_matrix = g["povm_matrix"]
        default, tight = [], []
        for c in g["counts"][:4]:
            tmg.results = c
            tmg.n_measurements = g["n_meas"]
            default.append(tmg.point_estimate("mle-constr").matrix)
            tight.append(tmg.point_estimate("mle-constr", tol=1e-12, max_iter=1000).matrix)
        out[tag + "_default"] = np.array(default)
        out[tag + "_tight"] = np.array(tight)
    return out


if __name__ == "__main__" and "--mle-constr" in sys.argv:
    warnings.filterwarnings("ignore")
    arrays = mle_constr_cases(load_reference())
    path = os.path.join(OUT, "mle_constr.npz")
    np.savez_compressed(path, **arrays)
    print(f"mle_constr: {os.path.getsize(path)} bytes")


def mhmc_cases(qp):
    """MHMCStateInterval (interval.py:689-759) with seeded legacy-RNG streams: inputs, the chain's samples and the
    sorted distances."""
    out = {}
    for tag, n, povm, shots, seed, kw in (
        ("q1", 1, "proj-set", 1000, 71, dict(n_points=120, step=0.05, burn_steps=60, thinning=2)),
        ("q2", 2, "proj", 10000, 72, dict(n_points=60, step=0.01, burn_steps=40, thinning=1)),
    ):
        rho = haar_mixed(n, np.random.default_rng(seed))
        tmg = qp.StateTomograph(qp.Qobj(rho))
        np.random.seed(seed)
        tmg.experiment(shots, povm)
        centre = tmg.point_estimate("mle")
        itv = qp.MHMCStateInterval(tmg, **kw)
        np.random.seed(seed + 1)
        itv.setup()
        out[tag + "_counts"] = tmg.results.copy()
        out[tag + "_povm"] = tmg.povm_matrix
        out[tag + "_n_meas"] = np.asarray(tmg.n_measurements, float)
        out[tag + "_centre"] = centre.matrix
        out[tag + "_seed"] = seed + 1
        out[tag + "_params"] = np.array([kw["n_points"], kw["step"], kw["burn_steps"], kw["thinning"]], float)
        out[tag + "_sorted_dist"] = itv.cl_to_dist.y
        out[tag + "_x_final"] = itv.chain.x_t
    return out


if __name__ == "__main__" and "--mhmc" in sys.argv:
    warnings.filterwarnings("ignore")
    arrays = mhmc_cases(load_reference())
    path = os.path.join(OUT, "mhmc.npz")
    np.savez_compressed(path, **arrays)
    print(f"mhmc: {os.path.getsize(path)} bytes")
