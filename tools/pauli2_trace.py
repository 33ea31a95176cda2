#!/usr/bin/env python
"""Per-warp timeline of k_mle_rrr_pauli2 (qpb_debug_set_trace): where a launch spends its time.
usage: pauli2_trace.py [B] [OPTION=value ...]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
from quantpy_b200 import _native as nt, engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
opts = dict(kv.split("=") for kv in sys.argv[2:])
rng = np.random.default_rng(0)
g = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
rho = g @ g.conj().T; rho /= np.trace(rho)
povm = qp.generate_measurement_matrix("proj", 2)
plan = engine.state_plan(povm, np.ones(1) * 10000)
probs = plan.probabilities(qp.Qobj(rho).bloch)[0]
lib = nt.load_library()
for k, v in opts.items():
    nt.set_option(k, int(v))
probs = probs.contiguous()
ref = nt.complex_to_device(rho)
bufs = plan.bootstrap_buffers(B)
iters = bufs["iters"]
def launch():  # the whole fused step: sampler, lin (+ start order), MLE with the hs distance
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plan.bootstrap_into(bufs, probs, ref, 1, 0, method="mle", max_iter=1000, tol=1e-6)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for _ in range(5):
    launch()
W = 148 * 12
trace = torch.zeros(W * 16 + 4 * B, dtype=torch.int64, device="cuda")
nt.check(lib.qpb_debug_set_trace(nt.ptr(trace), trace.numel() * 8))
ms = launch()
nt.check(lib.qpb_debug_set_trace(None, 0))
tall = trace.cpu().numpy()
t = tall[:W * 16].reshape(W, 16)
ps = tall[W * 16:].reshape(B, 4)
it = iters.cpu().numpy()
print(f"B={B} opts={opts}: step {ms:.3f} ms with trace; iterations mean {it.mean():.1f} max {it.max()} total {it.sum()}")
s = t[t[:, 0] > 0]            # thread-per-sample warps
w = t[(t[:, 8] > 0)]          # rows that ran the W worker (dedicated, or after their thread-per-sample phase)
t0 = min(s[:, 0].min(), w[:, 8].min() if len(w) else 1 << 62)
us = lambda x: (x - t0) / 1e3
print(f"thread-per-sample warps: {len(s)}; start spread {us(s[:,0]).max():.1f} us")
dr = s[s[:, 1] > 0]
print(f"  queue seen empty at   min {us(dr[:,1]).min():.0f}  median {np.median(us(dr[:,1])):.0f}  max {us(dr[:,1]).max():.0f} us; live lanes then: mean {dr[:,7].mean():.1f}")
print(f"  single phase ends at  min {us(s[:,2]).min():.0f}  p10 {np.quantile(us(s[:,2]),.1):.0f}  median {np.median(us(s[:,2])):.0f}  p90 {np.quantile(us(s[:,2]),.9):.0f}  max {us(s[:,2]).max():.0f} us")
wi, li, wit, lit = s[:, 3].sum(), s[:, 4].sum(), s[:, 5].sum(), s[:, 6].sum()
print(f"  warp-iterations {wi}  lane-iterations {li}  occupancy {li / wi / 32:.3f};  bulk: {wi - wit} / {li - lit} occ {(li - lit) / max(wi - wit, 1) / 32:.3f};  tail: {wit} / {lit} occ {lit / max(wit, 1) / 32:.3f}")
bulk_us = np.median(us(dr[:, 1]))
print(f"  bulk rate: {(li - lit) / bulk_us / 1e3:.2f} G sample-it/s over {bulk_us:.0f} us; per warp-iteration {bulk_us * len(s) / max(wi - wit, 1):.2f} us")
tail_span = (s[:, 2] - np.where(s[:, 1] > 0, s[:, 1], s[:, 2])) / 1e3
print(f"  tail per warp: mean {tail_span.mean():.0f} us, warp-iteration {tail_span.sum() / max(wit, 1):.2f} us")
print(f"  handed over {s[:,13].sum()} adopted {s[:,14].sum()}")
if len(w):
    print(f"W workers: {len(w)} rows; samples {w[:,10].sum()}  iterations {w[:,11].sum()}  busy {w[:,12].sum() / 1e3:.0f} us total -> {w[:,12].sum() / max(w[:,11].sum(), 1) / 1e3:.3f} us per iteration")
    print(f"  last exit {us(w[:,9]).max():.0f} us;  exits p10 {np.quantile(us(w[:,9]),.1):.0f} median {np.median(us(w[:,9])):.0f}")
    ded = w[w[:, 0] == 0]
    if len(ded):
        print(f"  dedicated: {len(ded)} warps, samples {ded[:,10].sum()}, iterations {ded[:,11].sum()}, busy share {ded[:,12].sum() / max((ded[:,9] - ded[:,8]).sum(), 1):.2f}")
# timeline: W iterations / single lane-iterations per 100 us cannot be reconstructed from totals; print histogram of single-phase ends
h, edges = np.histogram(us(s[:, 2]), bins=np.arange(0, us(s[:, 2]).max() + 100, 100))
print("  single-phase end histogram per 100 us:", h.tolist())

# per-sample timeline
st, pk, pu, fin = (us(ps[:, i].astype(np.float64)) for i in range(4))
parked = ps[:, 1] > 0
print(f"samples handed over at least once: {parked.sum()}; wait hand-over -> pick-up: median {np.median((pu - pk)[parked & (ps[:,2] > 0)]):.1f} max {((pu - pk)[parked & (ps[:,2] > 0)]).max():.1f} us")
last = np.argsort(fin)[-12:][::-1]
print("last finishers: (sample, iterations, start us, hand-over us, pick-up us, finish us)")
for b in last:
    print(f"   {b:6d} its {it[b]:4d}  start {st[b]:7.1f}  hand-over {pk[b] if ps[b,1] else -1:7.1f}  pick-up {pu[b] if ps[b,2] else -1:7.1f}  finish {fin[b]:7.1f}")
lng = np.argsort(it)[-8:][::-1]
print("longest samples:")
for b in lng:
    print(f"   {b:6d} its {it[b]:4d}  start {st[b]:7.1f}  hand-over {pk[b] if ps[b,1] else -1:7.1f}  pick-up {pu[b] if ps[b,2] else -1:7.1f}  finish {fin[b]:7.1f}")
for lo, hi in ((0, 100), (100, 200), (200, 300), (300, 450), (450, 600), (600, 1001)):
    m = (it >= lo) & (it < hi)
    if m.sum():
        print(f"   its [{lo},{hi}): n {m.sum():6d}  start median {np.median(st[m]):6.0f} p99 {np.quantile(st[m], .99):6.0f}  finish median {np.median(fin[m]):6.0f} max {fin[m].max():6.0f}")
h2, _ = np.histogram(fin, bins=np.arange(0, fin.max() + 100, 100))
print("   samples finished per 100 us:", h2.tolist())
