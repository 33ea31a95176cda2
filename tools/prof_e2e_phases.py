#!/usr/bin/env python
"""Where the end-to-end bootstrap call spends its time: phases of BootstrapStateInterval.setup timed with a
synchronize after each (adds a little, shows the split)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
from quantpy_b200 import engine, parallel
from quantpy_b200 import _native as nt
rng = np.random.default_rng(0)
g = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)); rho = g @ g.conj().T; rho /= np.trace(rho)
state = qp.Qobj(rho)
povm = qp.generate_measurement_matrix("proj", 2); n_meas = np.ones(1) * 10000
B = 100000
levels = np.linspace(1e-3, 1 - 1e-3, 1000)
def sync(): torch.cuda.synchronize()
acc = {}
def lap(name, t0):
    sync(); t = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t - t0); return t
for i in range(13):
    if i == 3: acc.clear()
    sync(); t = time.perf_counter()
    plan = engine.state_plan(povm, n_meas); t = lap("plan lookup (hash)", t)
    bloch = state.bloch; t = lap("state.bloch", t)
    probs = plan.probabilities(bloch)[0]; t = lap("probabilities (H2D + kernel)", t)
    bufs = plan.bootstrap_buffers(B); t = lap("buffers", t)
    ref = nt.complex_to_device(state.matrix); t = lap("ref H2D", t)
    plan.bootstrap_into(bufs, probs.contiguous(), ref, 100 + i, 0, "mle", True, "lin", 1000, 1e-6, "hs"); t = lap("kernels", t)
    q = parallel.quantile_function(bufs["dist"]); t = lap("sort", t)
    _ = q(levels); t = lap("quantile call (gather + D2H)", t)
tot = sum(acc.values())
for k, v in acc.items(): print(f"{k:34s} {v / 10 * 1e3:7.3f} ms per call")
print("sum per call %.3f ms" % (tot / 10 * 1e3))
tmg = qp.StateTomograph(state); tmg.povm_matrix, tmg.n_measurements = povm, n_meas
tmg.results = np.zeros((1, 36), dtype=np.int64); tmg.n_measurements = n_meas
sync(); t0 = time.perf_counter()
for i in range(10):
    itv = qp.BootstrapStateInterval(tmg, n_points=B, method="mle", tol=1e-6, max_iter=1000, state=state)
    itv.setup(seed=200 + i); _ = itv.cl_to_dist(levels)
sync(); print("public API per call %.3f ms" % ((time.perf_counter() - t0) * 100))
