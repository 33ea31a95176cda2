#!/usr/bin/env python
"""Fifth sweep: tail policy of k_mle_rrr_pauli2 (demand-driven hand-over, adoption) at C2; checks bit identity."""
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
def run(B, counts, start, out, iters, max_iter=1000, tol=1e-6, reps=5):
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nt.check(lib.qpb_mle_rrr(plan.handle, B, nt.ptr(counts), nt.ptr(start), max_iter, tol, nt.ptr(out), nt.ptr(iters), nt.stream_ptr()))
        e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return min(ms), float(np.median(ms))
class opts:
    def __init__(self, **kw): self.kw = kw; self.cm = []
    def __enter__(self):
        for k, v in self.kw.items():
            c = nt.option(k, v); c.__enter__(); self.cm.append(c)
    def __exit__(self, *e):
        for c in reversed(self.cm): c.__exit__(*e)
Bs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [12500, 25000, 50000, 100000]
for B in Bs:
    counts = plan.sample(probs, B, 1, 0)
    start = plan.lin(counts, True)
    out = torch.empty_like(start); iters = torch.empty(B, dtype=torch.int32, device="cuda")
    with opts(NO_TAIL_MERGE=1):
        tn, _ = run(B, counts, start, out, iters, reps=3)
    ref_out, ref_it = out.clone(), iters.clone()
    with opts(MLE_TAIL_POLL=-1):
        t_old, m_old = run(B, counts, start, out, iters)
    t_new, m_new = run(B, counts, start, out, iters)
    same = torch.equal(out, ref_out) and torch.equal(iters, ref_it)
    print(f"B={B}: thread-per-sample only {tn:.3f} ms, round-2a policy {t_old:.3f} (median {m_old:.3f}), new default {t_new:.3f} (median {m_new:.3f}) bit-identical {same}", flush=True)
    res = []
    for poll, age, adopt, live in itertools.product([2, 4, 8], [32, 64, 128, 200], [-1, 8, 32, 128], [3, 5, 8, 12, 16]):
        with opts(MLE_TAIL_POLL=poll, MLE_TAIL_AGE=age, MLE_ADOPT=adopt, MLE_PARK_LIVE=live):
            t, med = run(B, counts, start, out, iters, reps=3)
        ok = torch.equal(out, ref_out) and torch.equal(iters, ref_it)
        res.append((t, med, poll, age, adopt, live, ok))
    res.sort()
    print("   all bit-identical:", all(r[-1] for r in res))
    for r in res[:10]:
        print("   best  %.3f (med %.3f) poll %d age %d adopt %d live %d" % r[:6])
    for r in res[-3:]:
        print("   worst %.3f (med %.3f) poll %d age %d adopt %d live %d" % r[:6])
    # bulk hand-over age with the new tail
    for age in (200, 300, 450, 600, 1000000):
        with opts(MLE_PARK_AGE=age):
            t, med = run(B, counts, start, out, iters, reps=3)
        print(f"   park_age {age}: {t:.3f} (med {med:.3f})", flush=True)
    for ww in (-1, 2, 4):
        with opts(MLE_W_WARPS=ww):
            t, med = run(B, counts, start, out, iters, reps=3)
        print(f"   w_warps {ww}: {t:.3f} (med {med:.3f})", flush=True)
