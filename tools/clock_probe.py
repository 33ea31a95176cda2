#!/usr/bin/env python
"""Run the MLE kernel back to back and print per-launch times with the SM clock seen by nvidia-smi."""
import os, sys, subprocess, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantpy_b200 as qp
from quantpy_b200 import _native as nt, engine
B = 100000
rng = np.random.default_rng(0)
g = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)); rho = g @ g.conj().T; rho /= np.trace(rho)
povm = qp.generate_measurement_matrix("proj", 2)
plan = engine.state_plan(povm, np.ones(1) * 10000)
probs = plan.probabilities(qp.Qobj(rho).bloch)[0]
counts = plan.sample(probs, B, 1, 0); start = plan.lin(counts, True)
lib = nt.load_library(); out = torch.empty_like(start); iters = torch.empty(B, dtype=torch.int32, device="cuda")
rows = []
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader", "-lms", "20"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append((time.time(), l.strip())) for l in p.stdout], daemon=True).start()
time.sleep(0.3)
mi, tol = int(sys.argv[1]), float(sys.argv[2])
times = []
nosync = len(sys.argv) > 3
evs = []
for i in range(60):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nt.check(lib.qpb_mle_rrr(plan.handle, B, nt.ptr(counts), nt.ptr(start), mi, tol, nt.ptr(out), nt.ptr(iters), nt.stream_ptr()))
    e1.record()
    if nosync:
        evs.append((e0, e1))
    else:
        torch.cuda.synchronize(); times.append(e0.elapsed_time(e1))
if nosync:
    torch.cuda.synchronize(); times = [a.elapsed_time(b) for a, b in evs]
time.sleep(0.2); p.terminate()
print("times ms:", " ".join(f"{t:.2f}" for t in times))
print("smi:", " | ".join(r for _, r in rows[::3]))
