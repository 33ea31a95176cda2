#!/usr/bin/env python
"""cProfile of bench.py's e2e call (BootstrapStateInterval.setup + cl_to_dist) at C2."""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
_argv = sys.argv; sys.argv = sys.argv[:1]; args = bench.parse(); sys.argv = _argv
gpu = bench.Gpu(0, 0, 1)
cfg = bench.resolve_config(args, sys.argv[1] if len(sys.argv) > 1 else "c2")
sys.argv = sys.argv[:1]
wl = bench.make_workload(gpu, cfg)
levels = np.linspace(1e-3, 1 - 1e-3, 1000)
for i in range(3):
    itv = wl.interval(wl.B, seed=5 + i); _ = itv.cl_to_dist(levels)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10):
    itv = wl.interval(wl.B, seed=50 + i); _ = itv.cl_to_dist(levels)
torch.cuda.synchronize()
print("per call %.3f ms" % ((time.perf_counter() - t0) * 100))
pr = cProfile.Profile(); pr.enable()
for i in range(10):
    itv = wl.interval(wl.B, seed=80 + i); _ = itv.cl_to_dist(levels)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
