#!/usr/bin/env python
"""Profiling driver: sampler + lin projection at n qubits (default C3)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
from quantpy_b200 import engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
rng = np.random.default_rng(0); d = 2**n
g = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)); rho = g @ g.conj().T; rho /= np.trace(rho)
povm = qp.generate_measurement_matrix("proj", n)
plan = engine.state_plan(povm, np.ones(1) * 10000)
probs = plan.probabilities(qp.Qobj(rho).bloch)[0]
c = plan.sample(probs, B, 1, 0)
for rep in range(2):
    r = plan.lin(c, True)
torch.cuda.synchronize()
print("trace", r[..., 0].diagonal(dim1=1, dim2=2).sum(-1).mean().item())
