#!/usr/bin/env python
"""Time the counts GEMM (qpb_lin_project without projection = DMMA inversion + unpack) for the TMA pipeline and the
plain kernel.  Usage: python tools/prof_gemm.py [n_qubits] [B]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
from quantpy_b200 import _native as nt, engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
rng = np.random.default_rng(0)
d = 2**n
g = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)); rho = g @ g.conj().T; rho /= np.trace(rho)
pm = qp.generate_measurement_matrix("proj", n)
plan = engine.state_plan(pm, np.ones(1) * 10000)
probs = plan.probabilities(qp.Qobj(rho).bloch)[0]
counts = plan.sample(probs, B, 1, 0)
lib = nt.load_library()
out = torch.empty((B, d, d, 2), dtype=torch.float64, device="cuda")
def run(reps=7):
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nt.check(lib.qpb_lin_project(plan.handle, B, nt.ptr(counts), 0, nt.ptr(out), nt.stream_ptr()))
        e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return min(ms)
a = torch.randn(4096, 4096, device="cuda", dtype=torch.float64)
for _ in range(10):
    (a @ a).sum().item()
t = run()
with nt.option("NO_TMA_GEMM", 1):
    t0 = run()
fl = 2.0 * B * plan.K * plan.D
print(f"n={n} B={B} K={plan.K} D={plan.D}: inversion (GEMM + totals + unpack) TMA pipeline {t:.3f} ms, plain {t0:.3f} ms; GEMM flops {fl/1e9:.2f} G -> {fl/t/1e9:.1f} / {fl/t0/1e9:.1f} TFLOP/s incl. the other two kernels")
