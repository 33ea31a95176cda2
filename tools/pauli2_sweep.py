#!/usr/bin/env python
"""Time k_mle_rrr_pauli2 alone (CUDA events, min of several launches) for its lane mappings and hand-over policies.
Usage: python tools/pauli2_sweep.py [B ...]    -- C2 workload (2 qubits, 'proj', 1e4 shots, tol 1e-6, max_iter 1000)"""
import itertools, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
from quantpy_b200 import _native as nt, engine

sizes = [int(x) for x in sys.argv[1:]] or [100000]
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

def run(B, counts, start, out, iters, max_iter=1000, tol=1e-6, reps=7):
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nt.check(lib.qpb_mle_rrr(plan.handle, B, nt.ptr(counts), nt.ptr(start), max_iter, tol, nt.ptr(out), nt.ptr(iters), nt.stream_ptr()))
        e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return min(ms), float(np.median(ms))

for B in sizes:
    counts = plan.sample(probs, B, 1, 0)
    start = plan.lin(counts, True)
    out = torch.empty_like(start); iters = torch.empty(B, dtype=torch.int32, device="cuda")
    with nt.option("MLE_LANES", 1):
        t, med = run(B, counts, start, out, iters)
    it = iters.cpu().numpy()
    print(f"B={B}: mean its {it.mean():.1f} max {it.max()}  | thread-per-sample only: {t:.3f} ms (median {med:.3f})", flush=True)
    with nt.option("MLE_LANES", 1):
        tu, _ = run(B, counts, start, out, iters, max_iter=100, tol=0.0)
    print(f"   uniform 100 iterations, thread only: {tu:.3f} ms", flush=True)
    with nt.option("MLE_LANES", 32):
        t, med = run(B, counts, start, out, iters)
        tu, _ = run(B, counts, start, out, iters, max_iter=100, tol=0.0)
    print(f"   warp-per-sample only: {t:.3f} ms (median {med:.3f}); uniform 100 iterations: {tu:.3f} ms", flush=True)
    t, med = run(B, counts, start, out, iters)
    print(f"   default policy: {t:.3f} ms (median {med:.3f})", flush=True)
    if B >= 30000:
        for age, live, ww in itertools.product([100, 150, 200, 300, 400, 600], [4, 8, 16, 32], [-1, 2, 4]):
            with nt.option("MLE_LANES", 2), nt.option("MLE_PARK_AGE", age), nt.option("MLE_PARK_LIVE", live), nt.option("MLE_W_WARPS", ww):
                t, med = run(B, counts, start, out, iters, reps=4)
            print(f"   hybrid age {age:4d} live {live:3d} workers {ww:2d}: {t:.3f} ms (median {med:.3f})", flush=True)
