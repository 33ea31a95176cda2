import os, sys, time, cProfile, pstats
import numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import quantpy_b200 as qp
from quantpy_b200 import engine
chan = qp.channel.depolarizing(0.1, 2)
tmg = qp.ProcessTomograph(chan, "sic")
povm = qp.generate_measurement_matrix("proj-set", 2); n_meas = np.ones(povm.shape[0]) * 10000
tmg.adopt_measurement(povm, n_meas); tmg._process_plan()
centre = chan.choi.matrix
B = 1000
def step(i):
    counts = tmg.sample_counts(B, n_meas, povm, seed=i, offset=0, device=True)
    choi, iters = tmg.point_estimate_batch(counts, cptp=True, return_iters=True, device=True)
    return engine.distance(choi, centre, "hs")
for i in range(5): step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20): step(i)
torch.cuda.synchronize(); print("ms per step (back to back) %.3f" % ((time.perf_counter() - t0) * 50))
acc = {}
for i in range(20):
    torch.cuda.synchronize(); t = time.perf_counter()
    counts = tmg.sample_counts(B, n_meas, povm, seed=i, offset=0, device=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    choi, iters = tmg.point_estimate_batch(counts, cptp=True, return_iters=True, device=True); torch.cuda.synchronize(); t2 = time.perf_counter()
    d = engine.distance(choi, centre, "hs"); torch.cuda.synchronize(); t3 = time.perf_counter()
    for k, v in (("sample", t1 - t), ("estimate", t2 - t1), ("distance", t3 - t2)): acc[k] = acc.get(k, 0) + v
print({k: round(v * 50, 3) for k, v in acc.items()})
pr = cProfile.Profile(); pr.enable()
for i in range(50): step(i)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
