#!/usr/bin/env python
"""Eleventh sweep: hand-over ages at 1e5 / 5e4 samples after the W iteration became cheaper (212 instead of 248
instructions); whole fused bootstrap step, checks that distances and iteration counts keep their bits."""
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
def run(bufs, reps=5):
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
Bs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [100000, 50000]
for B in Bs:
    bufs = plan.bootstrap_buffers(B)
    with opts(NO_TAIL_MERGE=1, NO_MLE_ORDER=1):
        tn, _ = run(bufs, reps=3)
    ref_d, ref_it = bufs["dist"].clone(), bufs["iters"].clone()
    t_new, m_new = run(bufs, reps=8)
    same = torch.equal(bufs["dist"], ref_d) and torch.equal(bufs["iters"], ref_it)
    print(f"B={B}: thread-per-sample only {tn:.3f} ms, default {t_new:.3f} (median {m_new:.3f}) bit-identical {same}", flush=True)
    res = []
    grid = itertools.product([300, 350, 400, 450, 500, 600], [-1, 200, 250, 300, 350], [5, 8, 12], [-1, 4], [10, 25])
    for page, lo, live, poll, pct in grid:
        if lo > page: continue
        with opts(MLE_PARK_AGE=page, MLE_PARK_AGE_LO=lo, MLE_PARK_LIVE=live, MLE_TAIL_POLL=poll, MLE_PARK_AGE_PCT=pct):
            t, med = run(bufs, reps=4)
        ok = torch.equal(bufs["dist"], ref_d) and torch.equal(bufs["iters"], ref_it)
        res.append((t, med, page, lo, live, poll, pct, ok))
    res.sort(key=lambda r: r[1])
    print("   all bit-identical:", all(r[-1] for r in res))
    for r in res[:14]:
        print("   best  %.3f (med %.3f) park_age %d lo %d live %d poll %d pct %d" % r[:7])
    for r in res[-3:]:
        print("   worst %.3f (med %.3f) park_age %d lo %d live %d poll %d pct %d" % r[:7], flush=True)
