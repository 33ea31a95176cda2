import os, sys, time
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
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
def step(i, stamps=None):
    counts = tmg.sample_counts(B, n_meas, povm, seed=i, offset=0, device=True)
    if stamps is not None: stamps[0].record()
    choi, iters = tmg.point_estimate_batch(counts, cptp=True, return_iters=True, device=True)
    if stamps is not None: stamps[1].record()
    return engine.distance(choi, centre, "hs")
for i in range(5): step(i)
for do_flush in (False, True):
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(10)]
    for i, e in enumerate(evs):
        if do_flush: flush.zero_()
        e[0].record(); step(i, e[1:3]); e[3].record()
    torch.cuda.synchronize()
    a = np.array([[e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])] for e in evs])
    print("flush", do_flush, "mean ms: sample %.3f estimate %.3f distance %.3f total %.3f" % (*a.mean(0), a.sum(1).mean()))
