#!/usr/bin/env python
"""cProfile of the public-API bootstrap call at the reference's default size (1 qubit, n_points=1000)."""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
rng = np.random.default_rng(0)
g = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2)); rho = g @ g.conj().T; rho /= np.trace(rho)
state = qp.Qobj(rho)
tmg = qp.StateTomograph(state)
np.random.seed(0)
tmg.experiment(10000, "proj-set")
tmg.point_estimate("mle")
def call():
    itv = qp.BootstrapStateInterval(tmg, n_points=1000, method="mle", tol=1e-6, max_iter=1000)
    return itv()
for i in range(5): call()
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(50): call()
torch.cuda.synchronize(); print("ms per call %.3f" % ((time.perf_counter() - t0) * 20))
pr = cProfile.Profile(); pr.enable()
for i in range(50): call()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
