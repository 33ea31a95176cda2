#!/usr/bin/env python
"""A few end-to-end interval calls (BASELINE configs[1], then configs[0]) for an ncu launch list / a host profile:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv python tools/e2e_launches.py
   python tools/e2e_launches.py profile     (cProfile of the host side, sorted by cumulative time)"""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
levels = np.linspace(1e-3, 1 - 1e-3, 1000)
def make(n, povm, B):
    rng = np.random.default_rng(0); d = 2**n
    g = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)); rho = g @ g.conj().T; rho /= np.trace(rho)
    state = qp.Qobj(rho)
    tmg = qp.StateTomograph(state)
    tmg.povm_matrix = qp.generate_measurement_matrix(povm, n)
    tmg.results = np.zeros(tmg.povm_matrix.shape[:2], dtype=np.int64)
    tmg.n_measurements = np.ones(tmg.povm_matrix.shape[0]) * 10000
    def call(i):
        itv = qp.BootstrapStateInterval(tmg, n_points=B, method="mle", tol=1e-6, max_iter=1000, state=state)
        itv.setup(seed=200 + i)
        return itv.cl_to_dist(levels)
    return call
for n, povm, B in ((2, "proj", 100000), (1, "proj-set", 1000)):
    call = make(n, povm, B)
    for i in range(3): call(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(10): call(10 + i)
    torch.cuda.synchronize(); print(f"n={n} B={B}: {(time.perf_counter() - t0) * 100:.4f} ms per call", flush=True)
    if len(sys.argv) > 1 and sys.argv[1] == "profile":
        pr = cProfile.Profile(); pr.enable()
        for i in range(200 if B <= 1000 else 20): call(100 + i)
        pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
