#!/usr/bin/env python
"""How much would a start order buy the axis kernel at C4?  Run once, then feed the samples sorted by their own
iteration counts (longest first = the best possible order) and by the smallest positive eigenvalue of the unprojected
linear estimate (the predictor the two-qubit kernel uses)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
from quantpy_b200 import _native as nt, engine
lib = nt.load_library()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
rng = np.random.default_rng(0); d = 2**n
g = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)); rho = g @ g.conj().T; rho /= np.trace(rho)
povm = qp.generate_measurement_matrix("proj", n)
plan = engine.state_plan(povm, np.ones(1) * 10000)
probs = plan.probabilities(qp.Qobj(rho).bloch)[0]
counts = plan.sample(probs, B, 1234, 0).reshape(B, -1)
start = plan.lin(counts, True)
def run(c, s):
    out = torch.empty_like(s); iters = torch.empty(B, dtype=torch.int32, device="cuda")
    best = 1e9
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nt.check(lib.qpb_mle_rrr(plan.handle, B, nt.ptr(c), nt.ptr(s), 5000, 1e-6, nt.ptr(out), nt.ptr(iters), nt.stream_ptr()))
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best, iters
t0, it = run(counts, start)
itn = it.cpu().numpy()
print(f"n={n} B={B}: index order {t0:.2f} ms; iterations mean {itn.mean():.0f} p90 {np.quantile(itn,.9):.0f} max {itn.max()}")
perm = torch.from_numpy(np.argsort(-itn)).cuda()
t1, _ = run(counts[perm].contiguous(), start[perm].contiguous())
print(f"  longest first (oracle order): {t1:.2f} ms")
raw = plan.lin(counts, False).cpu().numpy(); raw = raw[..., 0] + 1j * raw[..., 1]
ev = np.linalg.eigvalsh(raw)
key = np.where(ev > 0, ev, np.inf).min(axis=1)
from scipy.stats import spearmanr
print(f"  spearman(iterations, smallest positive eigenvalue) = {spearmanr(key, itn)[0]:.3f}; negative eigenvalues per sample: mean {(ev < 0).sum(1).mean():.1f}")
perm2 = torch.from_numpy(np.argsort(key)).cuda()
t2, _ = run(counts[perm2].contiguous(), start[perm2].contiguous())
print(f"  ascending smallest positive eigenvalue: {t2:.2f} ms")
for name, k2 in (("sum of negative eigenvalues (most negative first)", np.where(ev < 0, ev, 0).sum(1)), ("number of positive eigenvalues (fewest first)", (ev > 0).sum(1) + 1e-3 * key),
                 ("second smallest positive", np.sort(np.where(ev > 0, ev, np.inf), axis=1)[:, 1])):
    print(f"  spearman(iterations, {name}) = {spearmanr(k2, itn)[0]:.3f}")
