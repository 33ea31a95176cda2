#!/usr/bin/env python
"""Profiling driver for the n-qubit axis MLE kernel (default n=4, 2000 samples, 50 iterations)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
from quantpy_b200 import engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
its = int(sys.argv[3]) if len(sys.argv) > 3 else 50
rng = np.random.default_rng(0); d = 2**n
g = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)); rho = g @ g.conj().T; rho /= np.trace(rho)
povm = qp.generate_measurement_matrix("proj", n)
plan = engine.state_plan(povm, np.ones(1) * 10000)
probs = plan.probabilities(qp.Qobj(rho).bloch)[0]
c = plan.sample(probs, B, 1, 0)
start = plan.lin(c, True)
for rep in range(2):
    r, it = plan.mle(c, start, its, 0.0)
torch.cuda.synchronize()
print("iters", it.double().mean().item())
