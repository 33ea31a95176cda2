#!/usr/bin/env python
"""Profiling driver: one sampler + lin + MLE + distance pass at BASELINE configs[1] size.
Usage: python tools/prof_mle.py [max_iter] [tol] [B]   (run under ncu -k regex:k_mle)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
from quantpy_b200 import _native as nt, engine

max_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-6
B = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
rng = np.random.default_rng(0)
g = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
rho = g @ g.conj().T; rho /= np.trace(rho)
povm = qp.generate_measurement_matrix("proj", 2)
plan = engine.state_plan(povm, np.ones(1) * 10000)
probs = plan.probabilities(qp.Qobj(rho).bloch)[0]
for rep in range(2):
    out = plan.bootstrap(probs, B, 1 + rep, 0, rho, method="mle", max_iter=max_iter, tol=tol)
torch.cuda.synchronize()
print("mean iters", out["iters"].double().mean().item(), "median dist", out["dist"].median().item())
if len(sys.argv) > 4 and sys.argv[4] == "time":   # MLE kernel(s) alone, CUDA events
    cnt = plan.sample(probs, B, 1, 0)
    start = plan.lin(cnt, True)
    best = 1e9
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = plan.mle(cnt, start, max_iter, tol); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("mle ms %.3f" % best)
