import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
from oracle import state as ostate
import quantpy_b200 as qp
B = 20000
rng = np.random.default_rng(0)
g = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
rho_t = g @ g.conj().T; rho_t /= np.trace(rho_t)
povm = qp.generate_measurement_matrix("proj", 2)
n_meas = np.ones(povm.shape[0]) * 10000
A = ostate.weighted_povm(povm, n_meas); E = ostate.povm_operators(A)
p = np.real(np.einsum("kab,ba->k", E, rho_t)).reshape(povm.shape[0], -1); p = p / p.sum(-1, keepdims=True)
counts = np.stack([rng.multinomial(10000, p[m], size=B) for m in range(p.shape[0])], axis=1)
rho0 = ostate.lin_estimate(counts, povm, n_meas, physical=True).astype(np.complex128)
f = counts.reshape(B, -1).astype(float); f /= f.sum(-1, keepdims=True)
def run(eps, seed=1):
    r = np.random.default_rng(seed)
    rho = rho0.copy(); iters = np.zeros(B, dtype=np.int32); active = np.arange(B)
    for it in range(1, 1001):
        if active.size == 0: break
        cur = rho[active]
        pk = np.real(np.einsum("kab,nba->nk", E, cur))
        w = f[active] / (pk + 1e-10)
        if eps: w = w * (1 + eps * (2 * r.random(w.shape) - 1))
        R = np.einsum("nk,kab->nab", w, E)
        new = R @ cur @ R
        new = 0.5 * (new + np.conj(np.swapaxes(new, -1, -2)))
        new /= np.real(np.trace(new, axis1=-2, axis2=-1))[:, None, None]
        d2 = np.sum(np.abs(new - cur) ** 2, axis=(-2, -1))
        rho[active] = new; iters[active] = it
        active = active[~(d2 < 1e-12)]
    return rho, iters
r0, i0 = run(0.0)
for eps in (2.0**-46, 2.0**-44):
    r1, i1 = run(eps)
    d = np.sqrt(np.sum(np.abs(r1 - r0) ** 2, axis=(-2, -1)))
    same = i1 == i0
    print(f"eps {eps:.2e}: iteration counts differ for {np.sum(~same)} of {B}; max Frobenius diff (equal counts) {d[same].max():.2e}; (all) {d.max():.2e}")
