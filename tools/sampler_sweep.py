#!/usr/bin/env python
"""Conditional-binomial sampler at BASELINE configs[1] size (1e5 x 36 outcomes x 1e4 shots): float32 prefilter of the
BTRS acceptance test on/off, period of the exact test, block size; prints ms per launch and a checksum of the counts
(which must not change)."""
import itertools, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
from quantpy_b200 import _native as nt, engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
Bs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [100000, 12500]
rng = np.random.default_rng(0); d = 2**n
g = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)); rho = g @ g.conj().T; rho /= np.trace(rho)
povm = qp.generate_measurement_matrix("proj", n)
plan = engine.state_plan(povm, np.ones(1) * 10000)
probs = plan.probabilities(qp.Qobj(rho).bloch)[0]
a = torch.randn(4096, 4096, device="cuda", dtype=torch.float64)
for _ in range(10):
    (a @ a).sum().item()
def run(B, reps=10):
    for rep in range(2):
        c = plan.sample(probs, B, 1 + rep, 0)
    torch.cuda.synchronize()
    ms = []
    for rep in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c = plan.sample(probs, B, 10, 0); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    w = torch.arange(1, c.numel() + 1, device="cuda", dtype=torch.int64)
    return min(ms), float(np.median(ms)), int((c.flatten().long() * w).sum().item())
with nt.option("SAMPLER", nt.SAMPLERS["binomial"]):
    for B in Bs:
        for groups, nopre, every, threads in itertools.product([1, 2, 3, 4, 6, 9, 12], [0], [1, 2], [0, 128, 64]):
            with nt.option("SAMPLER_LANES", groups), nt.option("SAMPLER_NO_PREFILTER", nopre), nt.option("SAMPLER_EXACT_EVERY", every), nt.option("SAMPLER_THREADS", threads):
                t, med, h = run(B)
            print(f"n {n} B {B} groups {groups} prefilter {1 - nopre} exact_every {every} threads {threads or 'auto'}: {t:.4f} ms (median {med:.4f}) checksum {h}", flush=True)
