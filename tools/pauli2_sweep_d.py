#!/usr/bin/env python
"""Fourth sweep: hand-over policy for batches below one wave of thread-per-sample lanes."""
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
for B in (1000, 3000, 5000, 8000, 12500, 25000, 37888, 50000, 75000, 100000):
    counts = plan.sample(probs, B, 1, 0)
    start = plan.lin(counts, True)
    out = torch.empty_like(start); iters = torch.empty(B, dtype=torch.int32, device="cuda")
    t0, _ = run(B, counts, start, out, iters)
    with nt.option("MLE_LANES", 32):
        tw, _ = run(B, counts, start, out, iters) if B <= 12500 else (float("nan"), 0)
    print(f"B={B}: default {t0:.3f} ms, W-only {tw:.3f}", flush=True)
    best = []
    for age, live in itertools.product([75, 100, 150, 200, 300, 450], [3, 5, 8, 16]):
        with nt.option("MLE_LANES", 2), nt.option("MLE_PARK_AGE", age), nt.option("MLE_PARK_LIVE", live):
            t, med = run(B, counts, start, out, iters, reps=4)
        best.append((t, age, live))
    best.sort()
    print("     best hybrid (ms, age, live):", " ".join(f"{t:.3f}/{a}/{l}" for t, a, l in best[:6]), " worst", f"{best[-1][0]:.3f}/{best[-1][1]}/{best[-1][2]}", flush=True)
