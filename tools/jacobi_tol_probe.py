#!/usr/bin/env python
"""Projection (lin, physical) at n = 3, 4: time and distance to the float128-free reference (numpy eigh) for the
library as built (QPB_JACOBI_REL2 is a compile-time constant of jacobi_rows.cu)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
from quantpy_b200 import engine
for n, B in ((3, 100000), (4, 10000)):
    rng = np.random.default_rng(0); d = 2**n
    g = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)); rho = g @ g.conj().T; rho /= np.trace(rho)
    povm = qp.generate_measurement_matrix("proj", n)
    plan = engine.state_plan(povm, np.ones(1) * 10000)
    probs = plan.probabilities(qp.Qobj(rho).bloch)[0]
    c = plan.sample(probs, B, 1, 0)
    raw = plan.lin(c, False); r = plan.lin(c, True)
    ms = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = plan.lin(c, True); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    m = min(4000, B)
    a = raw[:m].cpu().numpy(); a = a[..., 0] + 1j * a[..., 1]
    lam, v = np.linalg.eigh(a); lam = np.maximum(lam, 1e-15)
    want = np.einsum("bij,bj,bkj->bik", v, lam, v.conj()); want /= np.trace(want, axis1=1, axis2=2).real[:, None, None]
    got = r[:m].cpu().numpy(); got = got[..., 0] + 1j * got[..., 1]
    err = np.sqrt((np.abs(got - want) ** 2).sum((1, 2))).max()
    print(f"n={n} B={B}: lin + projection {min(ms):.3f} ms, max Frobenius error vs numpy eigh {err:.2e}")
