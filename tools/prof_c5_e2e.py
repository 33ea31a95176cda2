#!/usr/bin/env python
"""cProfile of BootstrapProcessInterval.setup() + cl_to_dist for a two-qubit channel (host side of config 5)."""
import os, sys, time, cProfile, pstats
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
chan = qp.channel.depolarizing(0.1, 2)
tmg = qp.ProcessTomograph(chan, "sic")
povm = qp.generate_measurement_matrix("proj-set", 2); n_meas = np.ones(povm.shape[0]) * 10000
tmg.adopt_measurement(povm, n_meas)
levels = np.linspace(1e-3, 1 - 1e-3, 1000)
def call(i):
    itv = qp.BootstrapProcessInterval(tmg, n_points=1000, method="lifp", cptp=True, channel=chan)
    itv.setup(seed=i)
    return itv.cl_to_dist(levels)
for i in range(3): call(i)
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(20): call(i)
torch.cuda.synchronize(); print("ms per call %.3f" % ((time.perf_counter() - t0) * 50))
pr = cProfile.Profile(); pr.enable()
for i in range(20): call(i)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
