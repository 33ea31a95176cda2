#!/usr/bin/env python
"""One short tiled-MLE call for an ncu launch list: prof_tiled_one.py n B iters"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
from quantpy_b200 import _native as nt, engine
lib = nt.load_library()
n, B, its = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(5); d = 2**n
g = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)); rho = g @ g.conj().T; rho /= np.trace(rho)
povm = qp.generate_measurement_matrix("sic", n)
plan = engine.state_plan(povm, np.ones(povm.shape[0]) * 10000)
probs = plan.probabilities(qp.Qobj(rho).bloch)[0]
counts = plan.sample(probs, B, 1, 0)
start = plan.lin(counts, True)
out = torch.empty_like(start); iters = torch.empty(B, dtype=torch.int32, device="cuda")
nt.check(lib.qpb_mle_rrr(plan.handle, B, nt.ptr(counts), nt.ptr(start), its, 0.0, nt.ptr(out), nt.ptr(iters), nt.stream_ptr()))
torch.cuda.synchronize()
