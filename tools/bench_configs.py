#!/usr/bin/env python
"""Time every BASELINE.json config (C1..C5) through the engine on one GPU (kernel pipeline only, CUDA events).
Usage: python tools/bench_configs.py [c1 c2 c3 c4 c5]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
from quantpy_b200 import _native as nt, engine

def haar(n, seed):
    rng = np.random.default_rng(seed); d = 2**n
    g = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)); r = g @ g.conj().T
    return r / np.trace(r)

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return min(ms), out

def state_case(name, n, povm, B, method, tol, max_iter):
    rho = haar(n, 0)
    pm = qp.generate_measurement_matrix(povm, n)
    plan = engine.state_plan(pm, np.ones(pm.shape[0]) * 10000)
    probs = plan.probabilities(qp.Qobj(rho).bloch)[0].contiguous()
    ref = nt.complex_to_device(rho); bufs = plan.bootstrap_buffers(B)
    ms, out = timed(lambda: plan.bootstrap_into(bufs, probs, ref, 1, 0, method=method, max_iter=max_iter, tol=tol))
    it = out["iters"].double().mean().item()
    t_s, _ = timed(lambda: plan.sample(probs, B, 1, 0))
    cnt = plan.sample(probs, B, 1, 0)
    t_l, start = timed(lambda: plan.lin(cnt, True))
    t_m = 0.0
    if method == "mle":
        t_m, _ = timed(lambda: plan.mle(cnt, start, max_iter, tol))
    print(f"{name}: n={n} povm={povm} K={plan.K} B={B} {method} tol={tol} max_iter={max_iter}: {ms:9.3f} ms  {B/ms/1e3:9.4f} Mrec/s  mean_it={it:.1f}"
          f"  [sampler {t_s:.3f} lin {t_l:.3f} mle {t_m:.3f} ms]", flush=True)

def process_case(name, n, B):
    chan = qp.channel.depolarizing(0.1, n)
    tmg = qp.ProcessTomograph(chan, "sic")
    np.random.seed(0); tmg.experiment(10000, "proj-set")
    centre = chan.choi.matrix
    def run():
        counts = tmg.sample_counts(B, 10000, "proj-set", seed=1, device=True)
        choi, iters = tmg.point_estimate_batch(counts, cptp=True, return_iters=True, device=True)
        return engine.distance(choi, centre, "hs"), iters
    ms, (d, iters) = timed(run, reps=2)
    print(f"{name}: process n={n} S={4**n} 'sic' inputs, proj-set, B={B} lifp+CPTP: {ms:9.3f} ms  {B/ms:9.2f} krec/s  mean CPTP iters={iters.double().mean().item():.1f}", flush=True)

def mhmc_case(n, povm, n_points, burn, chains):
    """MHMCStateInterval's chain (one warp per chain): the reference's single chain, and a batch with Philox noise."""
    rho = haar(n, 0)
    pm = qp.generate_measurement_matrix(povm, n)
    plan = engine.state_plan(pm, np.ones(pm.shape[0]) * 10000)
    probs = plan.probabilities(qp.Qobj(rho).bloch)[0].contiguous()
    counts = plan.sample(probs, 1, 1, 0)[0]
    x0 = engine.cholesky_vector(rho)
    rng = np.random.default_rng(1)
    total = burn + n_points
    dl, ul = rng.normal(size=(1, total, plan.D)), rng.random((1, total))
    t1, _ = timed(lambda: engine.mhmc_chains(plan, counts, x0[None], n_points, 0.01, burn, 1, dl, ul))
    tc, out = timed(lambda: engine.mhmc_chains(plan, counts, np.tile(x0, (chains, 1)), n_points, 0.01, burn, 1, seed=5))
    print(f"MHMC n={n} {povm}: 1 chain x {total} steps (host noise incl. H2D) {t1:.3f} ms = {t1 * 1e3 / total:.2f} us/step; "
          f"{chains} chains {tc:.3f} ms = {chains * total / tc / 1e3:.2f} Msteps/s, acceptance "
          f"{out['accepted'].double().mean().item() / n_points:.3f}", flush=True)


which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
if "c1" in which: state_case("C1", 1, "proj-set", 1000, "mle", 1e-6, 1000); state_case("C1b", 1, "proj-set", 100000, "mle", 1e-6, 1000)
if "c2" in which: state_case("C2", 2, "proj", 100000, "mle", 1e-6, 1000); state_case("C2-lin", 2, "proj", 100000, "lin", 0, 0)
if "c3" in which: state_case("C3", 3, "proj", 100000, "lin", 0, 0)
if "c3m" in which: state_case("C3-mle", 3, "proj", 2000, "mle", 1e-6, 1000)
if "c4" in which: state_case("C4-lin", 4, "proj", 10000, "lin", 0, 0); state_case("C4", 4, "proj", int(os.environ.get("C4B", "256")), "mle", 1e-6, 200)
if "mhmc" in which: mhmc_case(1, "proj-set", 1000, 1000, 4096); mhmc_case(2, "proj", 1000, 1000, 4096)
if "c5" in which: process_case("C5-1q", 1, 20000); process_case("C5-2q", 2, 500)
