#!/usr/bin/env python
"""Latency / throughput of the warp-per-sample mapping: uniform 200 iterations, W warps per SM = 1..12."""
import os, sys
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
for warps in (1, 2, 4, 8, 12):
    B = 148 * warps
    counts = plan.sample(probs, B, 1, 0)
    start = plan.lin(counts, True)
    out = torch.empty_like(start); iters = torch.empty(B, dtype=torch.int32, device="cuda")
    res = []
    for its in (200, 400):
        ms = []
        with nt.option("MLE_LANES", 32), nt.option("MLE_BLOCKS_PER_SM", warps):
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                nt.check(lib.qpb_mle_rrr(plan.handle, B, nt.ptr(counts), nt.ptr(start), its, 0.0, nt.ptr(out), nt.ptr(iters), nt.stream_ptr()))
                e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
        res.append(min(ms))
    per_it_us = (res[1] - res[0]) / 200 * 1e3
    print(f"{warps:2d} W warps per SM, one sample each: 200 its {res[0]:.3f} ms, 400 its {res[1]:.3f} ms -> {per_it_us:.3f} us = {per_it_us*1965:.0f} cycles per iteration", flush=True)
