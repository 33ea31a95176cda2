#!/usr/bin/env python
"""Twelfth sweep: hand-over age falling for the samples that start after the first wave of lanes (MLE_PARK_AGE_END /
_PCT2), re-measured after the W iteration became cheaper; whole fused bootstrap step at 1e5 samples."""
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
probs = plan.probabilities(qp.Qobj(rho).bloch)[0].contiguous()
ref = nt.complex_to_device(rho)
a = torch.randn(4096, 4096, device="cuda", dtype=torch.float64)
for _ in range(20):
    (a @ a).sum().item()
def run(bufs, reps=6):
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.bootstrap_into(bufs, probs, ref, 1, 0, method="mle", max_iter=1000, tol=1e-6)
        e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return min(ms), float(np.median(ms))
class opts:
    def __init__(self, **kw): self.kw = kw; self.cm = []
    def __enter__(self):
        for k, v in self.kw.items():
            c = nt.option(k, v); c.__enter__(); self.cm.append(c)
    def __exit__(self, *e):
        for c in reversed(self.cm): c.__exit__(*e)
B = 100000
bufs = plan.bootstrap_buffers(B)
with opts(NO_TAIL_MERGE=1, NO_MLE_ORDER=1):
    run(bufs, reps=2)
ref_d, ref_it = bufs["dist"].clone(), bufs["iters"].clone()
t, m = run(bufs, reps=10)
print(f"B={B}: default {t:.3f} (median {m:.3f})", flush=True)
res = []
for end, pct2, poll, live in itertools.product([150, 200, 250, 300, 350, 400], [10, 25, 50, 100], [-1, 4], [5, 8]):
    with opts(MLE_PARK_AGE_END=end, MLE_PARK_AGE_PCT2=pct2, MLE_TAIL_POLL=poll, MLE_PARK_LIVE=live):
        t, med = run(bufs)
    ok = torch.equal(bufs["dist"], ref_d) and torch.equal(bufs["iters"], ref_it)
    res.append((med, t, end, pct2, poll, live, ok))
res.sort()
print("   all bit-identical:", all(r[-1] for r in res))
for r in res[:12]:
    print("   best  median %.3f (min %.3f) age_end %d pct2 %d poll %d live %d" % r[:6])
for r in res[-3:]:
    print("   worst median %.3f (min %.3f) age_end %d pct2 %d poll %d live %d" % r[:6], flush=True)
