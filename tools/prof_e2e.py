#!/usr/bin/env python
"""cProfile of the public-API bootstrap call (host-side overheads of the e2e number)."""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
rng = np.random.default_rng(0)
g = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)); rho = g @ g.conj().T; rho /= np.trace(rho)
state = qp.Qobj(rho)
tmg = qp.StateTomograph(state)
tmg.povm_matrix = qp.generate_measurement_matrix("proj", 2)
tmg.results = np.zeros((1, 36), dtype=np.int64); tmg.n_measurements = np.ones(1) * 10000
def call(i):
    itv = qp.BootstrapStateInterval(tmg, n_points=100000, method="mle", tol=1e-6, max_iter=1000, state=state)
    itv.setup(seed=100 + i); return itv.cl_to_dist(0.95)
for i in range(3): call(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10): call(10 + i)
torch.cuda.synchronize()
print("ms per call", (time.perf_counter() - t0) * 100)
pr = cProfile.Profile(); pr.enable()
for i in range(10): call(30 + i)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
