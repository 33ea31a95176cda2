#!/usr/bin/env python
"""Iteration-count distribution of the R.rho.R MLE at BASELINE configs[1] (feeds the tail analysis in DESIGN.md)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
from quantpy_b200 import _native as nt, engine
n, B = 2, 100000
rng = np.random.default_rng(0); d = 2**n
g = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)); rho = g @ g.conj().T; rho /= np.trace(rho)
pm = qp.generate_measurement_matrix("proj", n)
plan = engine.state_plan(pm, np.ones(pm.shape[0]) * 10000)
probs = plan.probabilities(qp.Qobj(rho).bloch)[0].contiguous()
bufs = plan.bootstrap_buffers(B)
out = plan.bootstrap_into(bufs, probs, nt.complex_to_device(rho), 1, 0, method="mle", max_iter=1000, tol=1e-6)
it = out["iters"].cpu().numpy()
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/iters_c2.npy", it.astype(np.int16))
print("mean", it.mean(), "max", it.max(), "quantiles", np.quantile(it, [.5, .9, .99, .999]))
for cap in (100, 200, 300, 400, 500, 600):
    print(cap, "survivors", int((it > cap).sum()), "work after cap", int(np.clip(it - cap, 0, None).sum()))
