#!/usr/bin/env python
"""Tiled (DMMA GEMM) against warp-per-sample general-POVM MLE: 'sic' at n = 3, 4."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
from quantpy_b200 import _native as nt, engine
lib = nt.load_library()
def haar(n, seed):
    rng = np.random.default_rng(seed); d = 2**n
    g = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)); r = g @ g.conj().T
    return r / np.trace(r)
for n, B, max_iter, tol in ((3, 100000, 300, 1e-6), (3, 10000, 300, 1e-6), (4, 10000, 200, 1e-6), (4, 2000, 200, 1e-6)):
    rho = haar(n, 5)
    povm = qp.generate_measurement_matrix("sic", n)
    plan = engine.state_plan(povm, np.ones(povm.shape[0]) * 10000)
    probs = plan.probabilities(qp.Qobj(rho).bloch)[0]
    counts = plan.sample(probs, B, 1, 0)
    start = plan.lin(counts, True)
    out = torch.empty_like(start); iters = torch.empty(B, dtype=torch.int32, device="cuda")
    res = {}
    for name, off in (("tiled", 0), ("warp-per-sample", 1)):
        if off and n == 4 and B > 2000:
            continue
        with nt.option("NO_TILED_MLE", off):
            ms = []
            for _ in range(3):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                nt.check(lib.qpb_mle_rrr(plan.handle, B, nt.ptr(counts), nt.ptr(start), max_iter, tol, nt.ptr(out), nt.ptr(iters), nt.stream_ptr()))
                torch.cuda.synchronize(); ms.append((time.perf_counter() - t0) * 1e3)
            its = iters.cpu().numpy()
            res[name] = (min(ms), its.mean(), its.max(), out.clone())
    K, D = plan.K, 4**n
    line = f"n={n} K={K} B={B}: " + "; ".join(f"{k} {v[0]:.2f} ms (iterations mean {v[1]:.1f} max {v[2]})" for k, v in res.items())
    t = res["tiled"]
    flops = 4.0 * K * D * t[1] * B
    line += f"; tiled GEMM rate {flops / t[0] / 1e9:.2f} TFLOP/s of useful contraction work"
    if len(res) == 2:
        line += f"; max |diff| {float((res['tiled'][3] - res['warp-per-sample'][3]).abs().max()):.2e}"
    print(line, flush=True)
