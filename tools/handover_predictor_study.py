#!/usr/bin/env python
"""CPU study (oracle trajectories, no GPU): can the step norm predict, at a checkpoint every 32 iterations, whether a
sample of BASELINE configs[1] will still be running at iteration N?  del_k = ||rho_k - rho_{k-1}||_F^2 is recorded at
multiples of 32; the predictor extrapolates the geometric decay of the last 32 iterations."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import state as ostate
import quantpy_b200 as qp

B = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(0)
g = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
rho_t = g @ g.conj().T; rho_t /= np.trace(rho_t)
povm = qp.generate_measurement_matrix("proj", 2)
n_meas = np.ones(povm.shape[0]) * 10000
A = ostate.weighted_povm(povm, n_meas)
E = ostate.povm_operators(A)
p = np.real(np.einsum("kab,ba->k", E, rho_t)).reshape(povm.shape[0], -1)
p = p / p.sum(-1, keepdims=True)
counts = np.stack([rng.multinomial(10000, p[m], size=B) for m in range(p.shape[0])], axis=1)
lin_raw = ostate.lin_estimate(counts, povm, n_meas, physical=False)
lam_min = np.linalg.eigvalsh(lin_raw)[:, 0]
rho = ostate.lin_estimate(counts, povm, n_meas, physical=True).astype(np.complex128)
f = counts.reshape(B, -1).astype(float); f /= f.sum(-1, keepdims=True)
tol2 = 1e-12
iters = np.zeros(B, dtype=np.int32)
dels = np.full((B, 1000 // 32 + 1), np.nan)
active = np.arange(B)
for it in range(1, 1001):
    if active.size == 0: break
    cur = rho[active]
    pk = np.real(np.einsum("kab,nba->nk", E, cur))
    w = f[active] / (pk + 1e-10)
    R = np.einsum("nk,kab->nab", w, E)
    new = R @ cur @ R
    new = 0.5 * (new + np.conj(np.swapaxes(new, -1, -2)))
    new /= np.real(np.trace(new, axis1=-2, axis2=-1))[:, None, None]
    d2 = np.sum(np.abs(new - cur) ** 2, axis=(-2, -1))
    rho[active] = new
    iters[active] = it
    if it % 32 == 0: dels[active, it // 32] = d2
    active = active[~(d2 < tol2)]
np.savez("/tmp/study/traj.npz", iters=iters, dels=dels, lam_min=lam_min)
print("B", B, "mean its", iters.mean(), "p50/p90/p99/max", np.percentile(iters, [50, 90, 99]), iters.max())
for thr in (200, 250, 300, 350, 400, 450, 500):
    print(f"  P(n > {thr}) = {(iters > thr).mean():.4f}, mean excess {np.maximum(iters - thr, 0).sum() / B:.2f} its/sample")
