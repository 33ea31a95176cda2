#!/usr/bin/env python
"""Thirteenth sweep: the demand-driven tail (period of the look at the hand-over list, minimum age of the sample given
to a waiting W worker) at 1e5 and 5e4 samples; whole fused bootstrap step."""
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
for B in (100000, 50000):
    bufs = plan.bootstrap_buffers(B)
    with opts(NO_TAIL_MERGE=1, NO_MLE_ORDER=1):
        run(bufs, reps=2)
    ref_d, ref_it = bufs["dist"].clone(), bufs["iters"].clone()
    t, m = run(bufs, reps=10)
    print(f"B={B}: default {t:.3f} (median {m:.3f})", flush=True)
    res = []
    for poll, tage, live, adopt in itertools.product([1, 2, 4, 8, 16], [30, 60, 100, 150, 250], [5, 8], [-1, 4]):
        with opts(MLE_TAIL_POLL=poll, MLE_TAIL_AGE=tage, MLE_PARK_LIVE=live, MLE_ADOPT=adopt):
            t, med = run(bufs)
        ok = torch.equal(bufs["dist"], ref_d) and torch.equal(bufs["iters"], ref_it)
        res.append((med, t, poll, tage, live, adopt, ok))
    res.sort()
    print("   all bit-identical:", all(r[-1] for r in res))
    for r in res[:10]:
        print("   best  median %.3f (min %.3f) poll %d tail_age %d live %d adopt %d" % r[:6])
    for r in res[-3:]:
        print("   worst median %.3f (min %.3f) poll %d tail_age %d live %d adopt %d" % r[:6], flush=True)
