#!/usr/bin/env python
"""Phases of the strong-scaling step (BASELINE configs[1]: 1e5 resamples GLOBALLY over the ranks), CUDA events per phase,
max over ranks.  torchrun --nproc-per-node N tools/prof_strong_phases.py"""
import os, sys, ctypes
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
import quantpy_b200 as qp
from quantpy_b200 import _native as nt, engine, parallel as qpar
rng = np.random.default_rng(0)
g = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)); rho = g @ g.conj().T; rho /= np.trace(rho)
povm = qp.generate_measurement_matrix("proj", 2)
plan = engine.state_plan(povm, np.ones(1) * 10000)
probs = plan.probabilities(qp.Qobj(rho).bloch)[0].contiguous()
ref = nt.complex_to_device(rho)
N = 100000
lo, hi = qpar.shard_bounds(N, rank, world)
bufs = plan.bootstrap_buffers(hi - lo)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
names = ["bootstrap", "gather_sorted"]
acc = np.zeros(len(names))
steps = 20
def barrier():
    if world > 1:
        dist.barrier()
for i in range(steps + 3):
    flush.zero_(); barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    ev[0].record()
    plan.bootstrap_into(bufs, probs, ref, 100 + i, lo, method="mle", max_iter=1000, tol=1e-6)
    ev[1].record()
    full = qpar.gather_sorted(bufs["dist"], N)
    ev[2].record()
    torch.cuda.synchronize()
    if i >= 3:
        acc += [ev[k].elapsed_time(ev[k + 1]) for k in range(len(names))]
acc /= steps
t = torch.tensor(acc, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
# finer: sort, all-gather, merge separately (each its own events)
sub = np.zeros(3)
for i in range(steps):
    barrier()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record(); mine = engine.sort_f64(bufs["dist"]); e[1].record()
    if world > 1:
        gathered, width = qpar._all_gather_padded(mine, N)
    e[2].record()
    if world > 1:
        lens = np.zeros(world, dtype=np.int32); starts = np.zeros(world, dtype=np.int64)
        for r in range(world):
            l, h = qpar.shard_bounds(N, r, world); lens[r], starts[r] = h - l, r * width
        out = torch.empty((N,), dtype=torch.float64, device="cuda")
        nt.check(nt.load_library().qpb_merge_sorted_runs(world, lens.ctypes.data_as(ctypes.c_void_p), starts.ctypes.data_as(ctypes.c_void_p), nt.ptr(gathered), nt.ptr(out), nt.stream_ptr()))
    e[3].record(); torch.cuda.synchronize()
    sub += [e[k].elapsed_time(e[k + 1]) for k in range(3)]
sub /= steps
s = torch.tensor(sub, device="cuda")
if world > 1:
    dist.all_reduce(s, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world {world}: shard {hi - lo}: " + ", ".join(f"{n} {v:.3f} ms" for n, v in zip(names, t.tolist())) +
          f" | sort {s[0]:.3f}, all-gather {s[1]:.3f}, merge {s[2]:.3f} ms", flush=True)
if world > 1:
    dist.destroy_process_group()
