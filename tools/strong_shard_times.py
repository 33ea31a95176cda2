"""Per-shard step times of the strong-scaling record on ONE GPU: BASELINE configs[1] with 1e5 global resamples cut into
`world` shards, every shard's fused step (sampler -> lin -> MLE -> distance) timed alone with CUDA events for the seeds the
bench's timed steps use.  Separates a slow step that is a property of the samples (one shard's serial chain) from a
hiccup of the run (host, collective).

    python tools/strong_shard_times.py [world=8] [first_seed=1237] [n_seeds=10]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 1237
    n_seeds = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    sys.argv = sys.argv[:1]
    cfg = bench.resolve_config(bench.parse(), "c2")
    gpu = bench.Gpu(0, 0, 1)
    torch, qpar = gpu.torch, gpu.qpar
    n_total = cfg["resamples"]
    lo0, hi0 = qpar.shard_bounds(n_total, 0, world)
    wl = bench.StateWorkload(gpu, dict(cfg, resamples=hi0 - lo0))
    for i in range(5):
        wl.plan.bootstrap_into(wl.bufs, wl.probs, wl.ref, 1 + i, 0, **wl.kw)
    torch.cuda.synchronize()
    print(f"# ms per fused step of a {hi0 - lo0}-sample shard; rows = seed, columns = shard 0..{world - 1}; "
          "last columns: max over shards, largest iteration count of the seed")
    for seed in range(first, first + n_seeds):
        row, longest = [], 0
        for r in range(world):
            lo, hi = qpar.shard_bounds(n_total, r, world)
            best = 1e9
            for rep in range(3):
                gpu.flush.zero_()
                e0, e1 = gpu.event_pairs(1)[0]
                e0.record()
                wl.plan.bootstrap_into(wl.bufs, wl.probs, wl.ref, seed, lo, **wl.kw)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            longest = max(longest, int(wl.bufs["iters"].max().item()))
            row.append(best)
        print(seed, " ".join(f"{t:.3f}" for t in row), f"| {max(row):.3f} {longest}", flush=True)


if __name__ == "__main__":
    main()
