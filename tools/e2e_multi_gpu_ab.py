"""A/B of the multi-GPU end-to-end interval call inside ONE gpurun call (boxes differ by a few per cent):
A = the shard call followed by a stream synchronisation (the behaviour before round 2d: the host waited for the
shard before it queued the all-gather, the merge and the quantile call), B = the product path (the shard call
returns with its work queued; one synchronisation, in qpb_quantiles_host).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/e2e_multi_gpu_ab.py [calls]

Prints one line per variant: ms per call (wall clock per rank up to the call's return, MAX over ranks) and whether
both variants returned the same quantiles (they must: same seeds, same arithmetic)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    calls = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    gpu = bench.Gpu(rank, local_rank, world)
    torch, engine = gpu.torch, gpu.engine

    sys.argv = sys.argv[:1]                      # bench.parse() reads the command line: defaults only
    cfg = bench.resolve_config(bench.parse(), "c2")
    wl = bench.StateWorkload(gpu, cfg)
    levels = np.linspace(1e-3, 1 - 1e-3, 1000)
    plain = engine.bootstrap_interval

    def synced(*a, **k):
        out = plain(*a, **k)
        torch.cuda.current_stream().synchronize()
        return out

    results = {}
    for name, fn in (("A: synchronise after the shard call", synced), ("B: one synchronisation per call", plain),
                     ("A again", synced), ("B again", plain)):
        engine.bootstrap_interval = fn
        times, last = [], None
        bench.quiet_gc()
        for i in range(3 + calls):
            gpu.barrier()
            t0 = time.perf_counter()
            itv = wl.interval(wl.B * world, seed=wl.seed + 100 + i)
            last = itv.cl_to_dist(levels)
            t1 = time.perf_counter()
            gpu.barrier()
            if i >= 3:
                times.append(t1 - t0)
        total = gpu.max_over_ranks(float(np.sum(times)))
        results[name] = last
        if rank == 0:
            print(f"{name}: {1e3 * total / calls:.4f} ms per call, "
                  f"{world * wl.B * calls / total / 1e6:.1f} M rec/s on {world} GPU(s)", flush=True)
    engine.bootstrap_interval = plain
    keys = list(results)
    same = all(np.array_equal(results[keys[0]], results[k]) for k in keys[1:])
    if rank == 0:
        print("same quantiles from every variant:", same, flush=True)
    if world > 1:
        gpu.dist.barrier()
        gpu.dist.destroy_process_group()


if __name__ == "__main__":
    main()
