#!/usr/bin/env python
"""Stress of the pauli2 MLE scheduling: many fused bootstrap launches with random batch sizes, states, tolerances and
policy options back to back; every result is compared with the plain thread-per-sample launch of the same inputs.
A hang shows up as the timeout of the caller (run under `timeout`)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
from quantpy_b200 import _native as nt, engine
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_launch = int(sys.argv[2]) if len(sys.argv) > 2 else 200
povm = qp.generate_measurement_matrix("proj", 2)
plan = engine.state_plan(povm, np.ones(1) * 10000)
names = ["MLE_MERGE", "MLE_TAIL_POLL", "MLE_ADOPT", "MLE_PARK_LIVE", "MLE_PARK_AGE", "MLE_REFILL_MIN", "MLE_PARK_AGE_LO", "NO_MLE_ORDER", "MLE_W_WARPS"]
choices = {"MLE_MERGE": [0, 1, 4], "MLE_TAIL_POLL": [0, -1, 2, 8], "MLE_ADOPT": [0, 1, 16], "MLE_PARK_LIVE": [0, 1, 5, 31],
           "MLE_PARK_AGE": [0, 20, 150, 1000000], "MLE_REFILL_MIN": [0, 1, 3, 32], "MLE_PARK_AGE_LO": [0, -1, 30],
           "NO_MLE_ORDER": [0, 1], "MLE_W_WARPS": [0, -1, 2]}
bad = 0
t0 = time.time()
for i in range(n_launch):
    k = int(rng.integers(1, 5))
    g = rng.normal(size=(4, k)) + 1j * rng.normal(size=(4, k)); rho = g @ g.conj().T; rho /= np.trace(rho)
    probs = plan.probabilities(qp.Qobj(rho).bloch)[0].contiguous()
    ref = nt.complex_to_device(rho)
    B = int(rng.choice([1, 7, 33, 500, 3552, 3553, 4096, 9000, 37888, 37889, 60000, 100000, 150000]))
    tol = float(rng.choice([0.0, 1e-3, 1e-6])); max_iter = int(rng.choice([0, 1, 17, 200, 1000]))
    if tol == 0.0 and max_iter > 200: max_iter = 200
    bufs = plan.bootstrap_buffers(B)
    with nt.option("NO_TAIL_MERGE", 1), nt.option("NO_MLE_ORDER", 1):
        plan.bootstrap_into(bufs, probs, ref, 1000 + i, 0, method="mle", max_iter=max_iter, tol=tol)
    want_d, want_i = bufs["dist"].clone(), bufs["iters"].clone()
    opts = {nm: int(rng.choice(choices[nm])) for nm in names if rng.random() < 0.4}
    for nm, v in opts.items(): nt.set_option(nm, v)
    bufs["dist"].zero_(); bufs["iters"].fill_(-1)
    plan.bootstrap_into(bufs, probs, ref, 1000 + i, 0, method="mle", max_iter=max_iter, tol=tol)
    torch.cuda.synchronize()
    for nm in opts: nt.set_option(nm, 0)
    ok = torch.equal(bufs["dist"], want_d) and torch.equal(bufs["iters"], want_i)
    if not ok:
        bad += 1
        print("MISMATCH", i, B, tol, max_iter, k, opts, flush=True)
print(f"{n_launch} launches, {bad} mismatches, {time.time() - t0:.1f} s")
