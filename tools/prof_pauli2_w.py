#!/usr/bin/env python
"""One launch of the warp-per-sample mapping for ncu: python tools/prof_pauli2_w.py [warps_per_sm] [its] [samples_per_warp]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
from quantpy_b200 import _native as nt, engine
warps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
its = int(sys.argv[2]) if len(sys.argv) > 2 else 200
spw = int(sys.argv[3]) if len(sys.argv) > 3 else 1
rng = np.random.default_rng(0)
g = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
rho = g @ g.conj().T; rho /= np.trace(rho)
povm = qp.generate_measurement_matrix("proj", 2)
plan = engine.state_plan(povm, np.ones(1) * 10000)
probs = plan.probabilities(qp.Qobj(rho).bloch)[0]
lib = nt.load_library()
B = 148 * warps * spw
counts = plan.sample(probs, B, 1, 0)
start = plan.lin(counts, True)
out = torch.empty_like(start); iters = torch.empty(B, dtype=torch.int32, device="cuda")
with nt.option("MLE_LANES", 32), nt.option("MLE_BLOCKS_PER_SM", warps):
    for _ in range(2):
        nt.check(lib.qpb_mle_rrr(plan.handle, B, nt.ptr(counts), nt.ptr(start), its, 0.0, nt.ptr(out), nt.ptr(iters), nt.stream_ptr()))
torch.cuda.synchronize()
print("done", iters.float().mean().item())
