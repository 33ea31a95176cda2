#!/usr/bin/env python
"""Profiling driver for the multinomial sampler at BASELINE configs[1] size (and C3 with arg 3)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
from quantpy_b200 import engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
rng = np.random.default_rng(0); d = 2**n
g = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)); rho = g @ g.conj().T; rho /= np.trace(rho)
povm = qp.generate_measurement_matrix("proj", n)
plan = engine.state_plan(povm, np.ones(1) * 10000)
probs = plan.probabilities(qp.Qobj(rho).bloch)[0]
for rep in range(3):
    c = plan.sample(probs, B, 1 + rep, 0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for rep in range(reps):
    c = plan.sample(probs, B, 10 + rep, 0)
e1.record(); torch.cuda.synchronize()
print("n", n, "B", B, "sampler", os.environ.get("QPB_SAMPLER", "auto"), "flat" if os.environ.get("QPB_FLAT_BINOMIAL") else "grouped",
      "ms/launch %.3f" % (e0.elapsed_time(e1) / reps), "mean", c.double().mean().item())
