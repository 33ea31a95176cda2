#!/usr/bin/env python
"""Second sweep: W-only occupancy at small batches, hybrid hand-over policy at B = 1e5."""
import itertools, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
from quantpy_b200 import _native as nt, engine
rng = np.random.default_rng(0)
g = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
rho = g @ g.conj().T; rho /= np.trace(rho)
povm = qp.generate_measurement_matrix("proj", 2)
plan = engine.state_plan(povm, np.ones(1) * 10000)
probs = plan.probabilities(qp.Qobj(rho).bloch)[0]
lib = nt.load_library()
a = torch.randn(4096, 4096, device="cuda", dtype=torch.float64)
for _ in range(20):
    (a @ a).sum().item()
def run(B, counts, start, out, iters, max_iter=1000, tol=1e-6, reps=6):
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nt.check(lib.qpb_mle_rrr(plan.handle, B, nt.ptr(counts), nt.ptr(start), max_iter, tol, nt.ptr(out), nt.ptr(iters), nt.stream_ptr()))
        e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return min(ms), float(np.median(ms))
for B in (1000, 12500, 25000, 50000):
    counts = plan.sample(probs, B, 1, 0)
    start = plan.lin(counts, True)
    out = torch.empty_like(start); iters = torch.empty(B, dtype=torch.int32, device="cuda")
    with nt.option("MLE_LANES", 1):
        t1, _ = run(B, counts, start, out, iters)
    line = f"B={B}: thread only {t1:.3f} ms |"
    for ww in (16, 24, 32):
        with nt.option("MLE_LANES", 32), nt.option("MLE_W_WARPS", ww):
            t, _ = run(B, counts, start, out, iters)
        line += f" W x{ww}: {t:.3f}"
    with nt.option("MLE_LANES", 2):
        t, _ = run(B, counts, start, out, iters)
    line += f" | hybrid default {t:.3f}"
    print(line, flush=True)
B = 100000
counts = plan.sample(probs, B, 1, 0)
start = plan.lin(counts, True)
out = torch.empty_like(start); iters = torch.empty(B, dtype=torch.int32, device="cuda")
for age, live, ww in itertools.product([450, 600, 800, 100000], [1, 2, 3, 4, 6], [-1, 2, 4]):
    with nt.option("MLE_LANES", 2), nt.option("MLE_PARK_AGE", age), nt.option("MLE_PARK_LIVE", live), nt.option("MLE_W_WARPS", ww):
        t, med = run(B, counts, start, out, iters, reps=5)
    print(f"   hybrid age {age:6d} live {live:3d} workers {ww:2d}: {t:.3f} ms (median {med:.3f})", flush=True)
