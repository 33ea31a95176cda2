#!/usr/bin/env python
"""Benchmark of the tomography-bootstrap hot path (BASELINE.json metric:
"MLE bootstrap reconstructions/sec (n-qubit Pauli POVM)").

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--config c1|c2|c3|c4|c5-1q|c5-2q]

One *step* = one pass of the hot path over one batch per GPU: sample B count tables from the centre
object's POVM probabilities (Philox multinomial), reconstruct each and take its distance to the centre --
the body of quantpy/tomography/interval.py:598-609 (state) / :674-682 (process) for B = n_points resamples.

The headline workload is BASELINE.json configs[1] (`--config c2`, the default: 2 qubits, 36-outcome Pauli
'proj' POVM, 1e4 shots, 1e5 MLE resamples); under torchrun every rank processes its own 1e5 resamples of the
global index range (weak scaling, no data-path collective).  The JSON line also carries

`e2e`       the same metric through the public API (Bootstrap*Interval.setup() + cl_to_dist) with host inputs
            and host outputs inside the timed region (the all-gather + quantile step included).  For the state
            interval the API makes ONE C-ABI call per rank with host buffers on both sides
            (qpb_bootstrap_state_interval: upload, probabilities, bootstrap, sort, quantiles, download).  Every rank
            times its own calls (each returns after a stream synchronisation); the figure is the MAX over ranks.
            `e2e.dist_on_host` is the figure with all N sorted distances copied to the host as well (the
            reference's `cl_to_dist` owns them there).
`strong`    BASELINE configs[1] as written: 1e5 resamples GLOBALLY, sharded over the N ranks, with the sort of the
            shard, the all-gather and the merge of the sorted shards inside the CUDA-event-timed region.
`roofline`  for the dominant kernel of the config, timed alone: executed work / duration against a peak measured
            live in this run (FP64 FMA probe, shared-memory probe) -- see `roofline.bound`.
`cpu_baseline`  the oracle's port of the reference algorithm timed on one host core (bounded sample).
`configs`   (default run, one GPU) the other BASELINE configs through the same code, a few steps each.

--impl reference times the CPU port on all host cores and prints the same JSON shape.
"""

import argparse
import ctypes
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "reconstructions/s"

# BASELINE.json configs -> workloads.  tol / max_iter: R.rho.R step-norm threshold and cap (converged estimates);
# c4 is run to convergence as well (cap 5000, the histogram is in profiles/README_r2.md).
CONFIGS = {
    "c1": dict(index=0, kind="state", n_qubits=1, povm="proj-set", resamples=1000, method="mle", tol=1e-6, max_iter=1000),
    "c2": dict(index=1, kind="state", n_qubits=2, povm="proj", resamples=100000, method="mle", tol=1e-6, max_iter=1000),
    "c3": dict(index=2, kind="state", n_qubits=3, povm="proj", resamples=100000, method="lin", tol=0.0, max_iter=0),
    "c4": dict(index=3, kind="state", n_qubits=4, povm="proj", resamples=10000, method="mle", tol=1e-6, max_iter=5000),
    "c5-1q": dict(index=4, kind="process", n_qubits=1, povm="proj-set", resamples=10000, method="lifp", tol=1e-10,
                  max_iter=1000),
    "c5-2q": dict(index=4, kind="process", n_qubits=2, povm="proj-set", resamples=1000, method="lifp", tol=1e-10,
                  max_iter=1000),
}


def metric_name(cfg):
    if cfg["kind"] == "process":
        return f"process-tomography bootstrap reconstructions/sec ({cfg['n_qubits']}-qubit depolarising channel)"
    return f"{cfg['method'].upper()} bootstrap reconstructions/sec ({cfg['n_qubits']}-qubit Pauli POVM)"


def haar_mixed(n, seed, rank=0):
    """rho = G G^dagger / Tr, G complex Ginibre d x rank (rank 0 = full rank)."""
    rng = np.random.default_rng(seed)
    d = 2**n
    k = rank or d
    g = rng.normal(size=(d, k)) + 1j * rng.normal(size=(d, k))
    rho = g @ g.conj().T
    return rho / np.trace(rho)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--n-qubits", type=int, default=None)
    ap.add_argument("--povm", default=None)
    ap.add_argument("--shots", type=int, default=10000)
    ap.add_argument("--resamples", type=int, default=None, help="bootstrap resamples per GPU per step")
    ap.add_argument("--method", default=None, choices=["mle", "lin", "lifp", "states"])
    ap.add_argument("--tol", type=float, default=None, help="MLE step-norm stopping threshold")
    ap.add_argument("--max-iter", type=int, default=None, help="MLE iteration cap")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--state-seed", type=int, default=0)
    ap.add_argument("--state-rank", type=int, default=0, help="rank of the synthetic state (0 = full rank)")
    return ap.parse_args()


def resolve_config(args, name=None):
    cfg = dict(CONFIGS[name or args.config])
    cfg["name"] = name or args.config
    if name is None:  # command-line overrides apply to the main line only
        for key, val in (("n_qubits", args.n_qubits), ("povm", args.povm), ("resamples", args.resamples),
                         ("method", args.method), ("tol", args.tol), ("max_iter", args.max_iter)):
            if val is not None:
                cfg[key] = val
    cfg["shots"] = args.shots
    cfg["state_seed"], cfg["state_rank"] = args.state_seed, args.state_rank
    return cfg


def workload_config(cfg, world, impl="ours"):
    n = cfg["n_qubits"]
    outcomes = {"proj": str(6**n), "proj-set": f"{3**n}x{2**n}"}.get(cfg["povm"], "?")
    if cfg["kind"] == "process":
        what = (f"{n}-qubit depolarising channel (p = 0.1), 'sic' input states, {outcomes}-outcome Pauli '{cfg['povm']}' POVM, "
                f"{cfg['shots']} shots, {cfg['resamples']} '{cfg['method']}'+CPTP bootstrap resamples per GPU per step")
    else:
        rank = f", rank {cfg['state_rank']}" if cfg["state_rank"] else ""
        what = (f"{n}-qubit Haar-random state (seed {cfg['state_seed']}{rank}), {outcomes}-outcome Pauli '{cfg['povm']}' POVM, "
                f"{cfg['shots']} shots, {cfg['resamples']} {cfg['method'].upper()} bootstrap resamples per GPU per step")
    out = {"workload": f"BASELINE configs[{cfg['index']}]: {what}", "n_qubits": n, "povm": cfg["povm"], "shots": cfg["shots"],
           "resamples_per_gpu": cfg["resamples"], "global_resamples": cfg["resamples"] * world, "method": cfg["method"],
           "dst": "hs", "sharding": f"resamples x{world}"}
    if cfg["method"] == "mle":
        if impl == "reference":
            out.update(mle_update="SciPy BFGS over the Cholesky parametrisation (the reference's own 'mle', state.py:204-229)",
                       init="lin", tol=cfg["tol"], max_iter=cfg["max_iter"])
        else:
            out.update(mle_update="R.rho.R", init="lin", tol=cfg["tol"], max_iter=cfg["max_iter"])
    if cfg["kind"] == "process":
        out.update(cptp=True, cptp_tol=1e-12, cptp_max_iter=1000)
    if impl == "ours":
        out["l2"] = "flushed between timed steps (512 MiB write)"
    return out


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class NvmlSampler:
    """SM clock, power and throttle reasons polled through NVML every 2 ms from a thread of this process: the timed
    region of the default run is ~20 ms, too short for nvidia-smi's 100 ms loop to see more than once."""

    def __init__(self, index):
        import pynvml

        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        self.sm, self.power, self.mask = [], [], 0
        self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        self._stop = threading.Event()
        self.thread = None

    def start(self):
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        nv = self.nv
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                 ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                 ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "power_w_max": max(self.power) if self.power else None, "samples": len(self.sm),
                "reasons": sorted(n for n, bit in names if self.mask & bit), "source": "nvml, 2 ms poll"}


def clock_sampler(index):
    """NVML in-process when available, else an nvidia-smi loop."""
    try:
        return NvmlSampler(index)
    except Exception:
        return ClockSampler(index)


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); power.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference algorithm)
# ------------------------------------------------------------------------------------------------
def cpu_problem(cfg):
    """Host-side description of the workload for the oracle: (centre, povm tensor, shot vector, input states)."""
    from oracle import state as ostate

    n = cfg["n_qubits"]
    povm = ostate.measurement_matrix(cfg["povm"], n)
    n_meas = np.ones(povm.shape[0]) * cfg["shots"]
    if cfg["kind"] == "process":
        from oracle import process as oproc

        inputs = oproc.input_states("sic", n)
        return oproc.depolarizing_choi(0.1, n), povm, n_meas, inputs
    return haar_mixed(n, cfg["state_seed"], cfg["state_rank"]), povm, n_meas, None


def _cpu_chunk(seed, n_points, cfg, mle):
    import warnings

    from threadpoolctl import threadpool_limits

    from oracle import bootstrap as oboot

    warnings.filterwarnings("ignore")
    rng = np.random.RandomState(seed)
    centre, povm, n_meas, inputs = cpu_problem(cfg)
    with threadpool_limits(1):  # one BLAS thread per worker: the matrices are tiny, threads only contend
        t0 = time.perf_counter()
        if cfg["kind"] == "process":
            oboot.bootstrap_process(centre, inputs, povm, n_meas, n_points, method=cfg["method"], cptp=True, rng=rng,
                                    sort=False)
        else:
            oboot.bootstrap_state(centre, povm, n_meas, n_points, method=cfg["method"], tol=cfg["tol"],
                                  max_iter=cfg["max_iter"], mle=mle, rng=rng, sort=False)
        return time.perf_counter() - t0


def cpu_baseline(cfg, seconds):
    """Reference algorithm (oracle port: experiment -> estimate -> dst, serial loop) on ONE core, on a bounded
    sample of the same workload; for 'mle' also the NumPy R.rho.R port for an apples-to-apples figure."""
    import warnings

    warnings.filterwarnings("ignore")
    done, spent, chunk = 0, 0.0, 2
    while spent < seconds * 0.7:
        spent += _cpu_chunk(1000 + done, chunk, cfg, "bfgs")
        done += chunk
        if spent < seconds * 0.1:
            chunk = min(chunk * 2, 64)
    what = {"mle": "lin start + SciPy BFGS mle", "lin": "linear inversion + projection",
            "lifp": "lifp + CPTP projection", "states": "'states' assembly + CPTP projection"}[cfg["method"]]
    out = {"value": done / spent, "unit": UNIT, "cores": 1, "kind": "port",
           "sample": f"{done} resamples of the same workload through the oracle's serial bootstrap loop "
                     f"(reference algorithm: {what}), {spent:.1f} s on 1 core (the reference is single-threaded); "
                     f"host has {os.cpu_count()} cores"}
    if cfg["method"] == "mle" and cfg["n_qubits"] <= 2:
        from oracle import state as ostate
        from oracle.pauli import matrix_to_bloch

        centre, povm, n_meas, _ = cpu_problem(cfg)
        rng = np.random.RandomState(5)
        nb = 2000
        counts = ostate.experiment(povm, matrix_to_bloch(centre), n_meas, size=nb, rng=rng)
        t0 = time.perf_counter()
        ostate.mle_rrr(counts, povm, n_meas, max_iter=cfg["max_iter"], tol=cfg["tol"])
        dt = time.perf_counter() - t0
        out["rrr_numpy_port"] = {"value": nb / dt, "unit": UNIT, "cores": 1,
                                 "sample": f"{nb} resamples, vectorised NumPy R.rho.R oracle (same update and stopping "
                                           f"rule as the CUDA kernel), {dt:.1f} s"}
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port) on all host cores; rank 0 only."""
    if rank != 0:
        return
    import warnings
    from concurrent.futures import ProcessPoolExecutor

    warnings.filterwarnings("ignore")
    cfg = resolve_config(args)
    cores = os.cpu_count() or 1
    # size one step to ~3 s of wall time from a short probe
    probe = _cpu_chunk(1, 2, cfg, "bfgs") / 2
    per_core = max(1, min(400, int(3.0 / max(probe, 1e-4))))
    times = []
    with ProcessPoolExecutor(max_workers=cores) as pool:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            futs = [pool.submit(_cpu_chunk, 100 * step + c, per_core, cfg, "bfgs") for c in range(cores)]
            [f.result() for f in futs]
            if step >= args.warmup:
                times.append(time.perf_counter() - t0)
    total = float(np.sum(times))
    value = per_core * cores * args.steps / total
    sample = (f"each step = {per_core * cores} resamples ({per_core} per core x {cores} processes) of the same workload "
              f"through the oracle port of the reference algorithm")
    line = {"impl": "reference", "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg, world, impl="reference"), "gpu_launches": 0,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if cfg["method"] == "mle":
        # the same loop with the reference's own default stopping parameters (tol=1e-3, max_iter=100), 1 core
        dflt = dict(cfg, tol=1e-3, max_iter=100)
        t_def = _cpu_chunk(7, 20, dflt, "bfgs")
        line["api_defaults_1core"] = {"value": 20 / t_def, "unit": UNIT, "tol": 1e-3, "max_iter": 100,
                                      "note": "reference algorithm with StateTomograph.point_estimate's default tol/max_iter"}
    print(json.dumps(line), flush=True)


def kernel_models():
    """Per-kernel constants taken from committed ncu captures (profiles/kernel_models.json): executed FP64
    instructions or shared-memory wavefronts per unit of work, DRAM bytes per launch."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "kernel_models.json")))
    except OSError:
        return {}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def quiet_gc():
    """Collect now and keep the cyclic collector out of a timed region (a generation-2 pass over the objects the CPU
    baseline leaves behind stalled one step of ten by 13 ms); the caller re-enables it."""
    gc.collect()
    gc.disable()


class Gpu:
    """Process-wide handles of the GPU arm."""

    def __init__(self, rank, local_rank, world):
        import torch
        import torch.distributed as dist

        import quantpy_b200 as qp
        from quantpy_b200 import _native as nt
        from quantpy_b200 import engine
        from quantpy_b200 import parallel as qpar

        self.torch, self.dist, self.qp, self.nt, self.engine, self.qpar = torch, dist, qp, nt, engine, qpar
        self.rank, self.local_rank, self.world = rank, local_rank, world
        torch.cuda.set_device(local_rank)
        if world > 1 and not dist.is_initialized():
            # stdout carries exactly one JSON line: NCCL prints its banner ("NCCL version ...") to fd 1 when the
            # communicator comes up, so fd 1 points at stderr until the first collective has run
            sys.stdout.flush()
            saved_stdout = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
                dist.barrier()
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved_stdout, 1)
                os.close(saved_stdout)
        self.lib = nt.load_library()
        self.flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
        self._fp64_peak = None
        self._smem_peak = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def event_pairs(self, n):
        ev = self.torch.cuda.Event
        return [(ev(enable_timing=True), ev(enable_timing=True)) for _ in range(n)]

    def time_alone(self, fn, reps=5, flush=True):
        """Mean CUDA-event duration of fn() launched alone (L2 flushed before each launch)."""
        pairs = self.event_pairs(reps)
        fn()
        self.torch.cuda.synchronize()
        for e0, e1 in pairs:
            if flush:
                self.flush.zero_()
            e0.record()
            fn()
            e1.record()
        self.torch.cuda.synchronize()
        return float(np.mean([e0.elapsed_time(e1) for e0, e1 in pairs]))

    def fp64_peak(self):
        """FP64 FMA peak (TFLOP/s), measured now on this device (MEASURED_PEAKS.json has no FP64 line)."""
        if self._fp64_peak is None:
            torch, nt = self.torch, self.nt
            sink = torch.zeros(8, dtype=torch.float64, device="cuda")
            flops = np.zeros(1)
            best = 0.0
            for _ in range(3):
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                p0.record()
                nt.check(self.lib.qpb_fp64_fma_probe(200000, nt.ptr(sink), flops.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                                     nt.stream_ptr()))
                p1.record()
                torch.cuda.synchronize()
                best = max(best, flops[0] / (p0.elapsed_time(p1) * 1e-3) / 1e12)
            self._fp64_peak = best
        return self._fp64_peak

    def smem_peak(self):
        """Shared-memory wavefronts per second over the whole device, measured now (conflict-free 64-bit loads)."""
        if self._smem_peak is None:
            torch, nt = self.torch, self.nt
            sink = torch.zeros(8, dtype=torch.float64, device="cuda")
            wf = np.zeros(1)
            best = 0.0
            for _ in range(3):
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                p0.record()
                nt.check(self.lib.qpb_smem_probe(20000, nt.ptr(sink), wf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                                 nt.stream_ptr()))
                p1.record()
                torch.cuda.synchronize()
                best = max(best, wf[0] / (p0.elapsed_time(p1) * 1e-3))
            self._smem_peak = best
        return self._smem_peak


def hbm_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return peaks["hbm_gbs"], "MEASURED_PEAKS.json"
    except (OSError, KeyError):
        return 6650.0, "fallback of B200_PROFILING.md"


class StateWorkload:
    """configs c1..c4: BootstrapStateInterval's loop."""

    def __init__(self, gpu, cfg):
        qp, nt, engine = gpu.qp, gpu.nt, gpu.engine
        self.gpu, self.cfg = gpu, cfg
        n, self.B = cfg["n_qubits"], cfg["resamples"]
        self.centre = haar_mixed(n, cfg["state_seed"], cfg["state_rank"])
        self.state = qp.Qobj(self.centre)
        self.povm = qp.generate_measurement_matrix(cfg["povm"], n)
        self.n_meas = np.ones(self.povm.shape[0]) * cfg["shots"]
        self.plan = engine.state_plan(self.povm, self.n_meas)
        self.probs = self.plan.probabilities(self.state.bloch)[0].contiguous()   # resident inputs
        self.ref = nt.complex_to_device(self.centre)
        self.bufs = self.plan.bootstrap_buffers(self.B)
        self.kw = dict(method=cfg["method"], physical=True, init="lin", max_iter=cfg["max_iter"], tol=cfg["tol"], dst="hs")
        self.seed = 1234

    def step(self, i):
        # Philox counter = global sample index: rank r owns [r*B, (r+1)*B) of step i's range
        self.plan.bootstrap_into(self.bufs, self.probs, self.ref, self.seed + i, self.gpu.rank * self.B, **self.kw)

    def mean_iterations(self):
        return float(self.bufs["iters"].double().mean().item()) if self.cfg["method"] == "mle" else None

    def interval(self, n_points, seed):
        qp = self.gpu.qp
        if not hasattr(self, "tmg"):
            self.tmg = qp.StateTomograph(self.state)
            self.tmg.povm_matrix = self.povm
            self.tmg.results = np.zeros(self.povm.shape[:2], dtype=np.int64)  # bookkeeping only; the centre is given explicitly
            self.tmg.n_measurements = self.n_meas
        itv = qp.BootstrapStateInterval(self.tmg, n_points=n_points, method=self.cfg["method"], tol=self.cfg["tol"],
                                        max_iter=self.cfg["max_iter"], state=self.state)
        itv.setup(seed=seed)
        return itv

    def e2e_h2d(self):
        if self.gpu.engine.FUSED_INTERVAL:
            return 0  # one library call uploads Bloch vector, centre state and levels; counted in parallel.TRAFFIC
        return self.state.bloch.size * 8 + self.ref.numel() * 8

    api = ("quantpy_b200.BootstrapStateInterval(...).setup() + cl_to_dist; on one GPU both are served by ONE C-ABI call "
           "with host inputs and outputs (qpb_bootstrap_state_interval)")

    # ---- dominant kernel, timed alone ---------------------------------------------------------------
    def kernel_breakdown(self):
        gpu, plan, B = self.gpu, self.plan, self.B
        torch, nt, lib = gpu.torch, gpu.nt, gpu.lib
        out = {}
        out["sampler"] = gpu.time_alone(lambda: plan.sample(self.probs, B, 1, 0))
        counts = plan.sample(self.probs, B, 1, 0)
        rho_lin = torch.empty((B, plan.d, plan.d, 2), dtype=torch.float64, device="cuda")
        out["lin_project"] = gpu.time_alone(lambda: nt.check(lib.qpb_lin_project(
            plan.handle, B, nt.ptr(counts), 1, nt.ptr(rho_lin), nt.stream_ptr())))
        if plan.n_qubits >= 3:
            out["lin_inversion_only"] = gpu.time_alone(lambda: nt.check(lib.qpb_lin_project(
                plan.handle, B, nt.ptr(counts), 0, nt.ptr(rho_lin), nt.stream_ptr())))
        self._counts, self._start = counts, rho_lin
        if self.cfg["method"] == "mle":
            rho = torch.empty_like(rho_lin)
            self._iters = torch.empty((B,), dtype=torch.int32, device="cuda")
            # the fused call starts the likely long runners first (qpb_lin_project_ordered); time the kernel the same way
            order = torch.empty((B,), dtype=torch.int32, device="cuda")
            nt.check(lib.qpb_lin_project_ordered(plan.handle, B, nt.ptr(counts), nt.ptr(rho_lin), nt.ptr(order),
                                                 nt.stream_ptr()))
            out["mle"] = gpu.time_alone(lambda: nt.check(lib.qpb_mle_rrr_ordered(
                plan.handle, B, nt.ptr(counts), nt.ptr(rho_lin), nt.ptr(order), self.cfg["max_iter"], self.cfg["tol"],
                nt.ptr(rho), nt.ptr(self._iters), nt.stream_ptr())))
            if int(lib.qpb_mle_variant(plan.handle)) == 3:
                out["mle_index_order"] = gpu.time_alone(lambda: nt.check(lib.qpb_mle_rrr(
                    plan.handle, B, nt.ptr(counts), nt.ptr(rho_lin), self.cfg["max_iter"], self.cfg["tol"], nt.ptr(rho),
                    nt.ptr(self._iters), nt.stream_ptr())))
        return out

    def roofline(self, kernels):
        gpu, plan, B, cfg = self.gpu, self.plan, self.B, self.cfg
        models = kernel_models()
        K, D, d = plan.K, plan.D, plan.d
        hbm, hbm_src = hbm_peak()
        if cfg["method"] == "mle":
            k_ms = kernels["mle"]
            it = self._iters.double()
            tot_iters = float(it.sum().item())
            variant = {0: "k_mle_rrr_generic", 1: "k_mle_rrr_small", 2: "k_mle_rrr_const", 3: "k_mle_rrr_pauli2",
                       4: "k_mle_rrr_axis", 5: "k_mle_rrr_tiled"}[int(gpu.lib.qpb_mle_variant(plan.handle))]
            dense = 4 * K * D + 16 * d**3                                # SURVEY section 8d, dense contraction
            alg_bytes = B * (4 * K + 2 * 16 * D + 4)  # counts in, start state in, state out, iteration count out
            base = {"kernel": variant, "kernel_ms": k_ms, "mean_iterations": tot_iters / B,
                    "max_iterations": int(it.max().item()),
                    "traffic": models.get(variant, {}).get("dram_bytes_per_launch"),
                    "hbm": {"achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                            "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm, "peak_source": hbm_src,
                            "note": "structurally tiny: the iteration never touches HBM"},
                    "survey_8d": {"flop_per_iteration": dense, "achieved": tot_iters * dense / (k_ms * 1e-3) / 1e12,
                                  "note": "dense-equivalent rate (algorithmic flops of SURVEY 8d / kernel time); NOT a "
                                          "pipe utilisation for the structured kernels, which execute fewer flops"}}
            if variant == "k_mle_rrr_axis":
                # bound by shared-memory wavefronts (ncu: >90 % of the LSU data pipe); executed flops are a small
                # fraction of the dense count, so an FP64 fraction would say nothing
                wf_iter = models.get(variant, {}).get(f"wavefronts_per_sample_iteration_n{plan.n_qubits}")
                peak = gpu.smem_peak()
                if wf_iter:
                    ach = tot_iters * wf_iter / (k_ms * 1e-3)
                    base.update(bound="smem", achieved=ach / 1e9, peak=peak / 1e9, unit="Gwavefront/s", frac=ach / peak,
                                work_note=f"{wf_iter} shared-memory wavefronts per sample-iteration (ncu l1tex data-pipe "
                                          "wavefronts / sample-iterations, profiles/kernel_models.json)",
                                peak_source="qpb_smem_probe, measured in this run (conflict-free LDS.64 from every SM)")
                else:
                    base.update(bound="smem", achieved=None, peak=peak / 1e9, unit="Gwavefront/s", frac=None,
                                work_note="no wavefront count committed for this kernel")
                return base
            peak = gpu.fp64_peak()
            m = models.get(variant, {})
            if "fp64_instructions_per_sample_iteration" in m:
                inst, fma = m["fp64_instructions_per_sample_iteration"], m["dfma_per_sample_iteration"]
                flop_iter = inst + fma  # FMA = 2 flop, add / mul = 1
                note = (f"executed: {fma} DFMA + {inst - fma} DADD/DMUL per sample-iteration of the thread-per-sample "
                        "mapping (ncu thread-instruction counts / sample-iterations, profiles/kernel_models.json)")
                pipe = tot_iters * inst * 2.0 / (k_ms * 1e-3) / 1e12 / peak
            else:
                flop_iter, note, pipe = dense, "dense kernel: SURVEY 8d's 4KD + 16 d^3 flop per iteration are executed", None
            ach = tot_iters * flop_iter / (k_ms * 1e-3) / 1e12
            base.update(bound="fp64", achieved=ach, peak=peak, unit="TFLOP/s", frac=ach / peak, flop_per_iteration=flop_iter,
                        work_note=note, fp64_pipe_frac=pipe,
                        peak_source="qpb_fp64_fma_probe, measured in this run (MEASURED_PEAKS.json has no FP64 line)")
            return base
        # 'lin': the sampler dominates at n = 3 (alias draws: one per shot); report it against the shared-memory
        # roofline and the DMMA inversion against the FP64 one
        out = {"kernel": "k_multinomial (alias)" if kernels["sampler"] >= kernels["lin_project"] else "lin_project",
               "kernel_ms": max(kernels["sampler"], kernels["lin_project"])}
        shots = B * plan.P * cfg["shots"]
        m = models.get("k_multinomial", {})
        peak = gpu.smem_peak()
        if kernels["sampler"] >= kernels["lin_project"] and "wavefronts_per_draw" in m:
            ach = shots * m["wavefronts_per_draw"] / (kernels["sampler"] * 1e-3)
            out.update(bound="smem", achieved=ach / 1e9, peak=peak / 1e9, unit="Gwavefront/s", frac=ach / peak,
                       work_note=f"{m['wavefronts_per_draw']} shared-memory wavefronts per categorical draw (ncu, "
                                 "profiles/kernel_models.json), one draw per shot",
                       peak_source="qpb_smem_probe, measured in this run")
        elif kernels["sampler"] >= kernels["lin_project"]:
            out.update(kernel="k_multinomial_binomial (two passes)", bound="issue", achieved=None, peak=None, unit=None, frac=None,
                       work_note="conditional-binomial sampler: bound by fixed-latency dependencies and divergence of the per-lane "
                                 "state machine (profiles/ncu_r2c_btrs.txt: FP64 pipe 12 %, 17 of 32 lanes active), no pipe roofline")
        else:
            pr = models.get("k_project_rows", {})
            key = f"flop_per_matrix_d{plan.d}"
            if key in pr and "lin_inversion_only" in kernels:
                j_ms = kernels["lin_project"] - kernels["lin_inversion_only"]
                ach = B * pr[key] / (j_ms * 1e-3) / 1e12
                pk = gpu.fp64_peak()
                out.update(kernel="k_project_rows (register-resident Jacobi + projection)", kernel_ms=j_ms, bound="fp64",
                           achieved=ach, peak=pk, unit="TFLOP/s", frac=ach / pk,
                           fp64_pipe_frac=B * pr[f"fp64_instructions_per_matrix_d{plan.d}"] * 2.0 / (j_ms * 1e-3) / 1e12 / pk,
                           work_note=f"executed: {pr[key]} flop per matrix (ncu thread-instruction counts, "
                                     "profiles/kernel_models.json); kernel time = lin_project - DMMA inversion, both timed alone; "
                                     "bound by fixed-latency dependencies of the rotation chain (profiles/ncu_r2c_rows8.txt)",
                           peak_source="qpb_fp64_fma_probe, measured in this run")
            else:
                out.update(kernel="lin_project (k_gemm_counts_dmma + k_project_rows)", bound="issue", achieved=None, peak=None,
                           unit=None, frac=None,
                           work_note="the register-resident Jacobi projection dominates; no instruction count committed for this size")
        if "lin_inversion_only" in kernels:
            g_ms = kernels["lin_inversion_only"]
            fl = 2.0 * B * K * D
            out["dmma_inversion"] = {"kernel": "k_gemm_counts_dmma (+ unpack)", "kernel_ms": g_ms, "bound": "fp64 tensor",
                                     "achieved": fl / (g_ms * 1e-3) / 1e12, "peak": gpu.fp64_peak(), "unit": "TFLOP/s",
                                     "frac": fl / (g_ms * 1e-3) / 1e12 / gpu.fp64_peak(),
                                     "work_note": "2 K D flop per sample; DMMA and DFMA share the FP64 pipe (16 FMA/clk/SMSP)"}
        return out


class ProcessWorkload:
    """config c5: BootstrapProcessInterval's loop ('lifp' + CPTP projection)."""

    def __init__(self, gpu, cfg):
        qp = gpu.qp
        self.gpu, self.cfg = gpu, cfg
        n, self.B = cfg["n_qubits"], cfg["resamples"]
        self.chan = qp.channel.depolarizing(0.1, n)
        self.tmg = qp.ProcessTomograph(self.chan, "sic")
        self.povm = qp.generate_measurement_matrix(cfg["povm"], n)
        self.n_meas = np.ones(self.povm.shape[0]) * cfg["shots"]
        self.tmg.adopt_measurement(self.povm, self.n_meas)
        self.centre = self.chan.choi.matrix
        self.seed = 1234
        self.tmg._process_plan()
        self.last_iters = None

    def step(self, i):
        gpu, B = self.gpu, self.B
        counts = self.tmg.sample_counts(B, self.n_meas, self.povm, seed=self.seed + i, offset=gpu.rank * B, device=True)
        choi, iters = self.tmg.point_estimate_batch(counts, cptp=True, return_iters=True, device=True)
        self.last = gpu.engine.distance(choi, self.centre, "hs")
        self.last_iters = iters

    def mean_iterations(self):
        return float(self.last_iters.double().mean().item())

    def interval(self, n_points, seed):
        qp = self.gpu.qp
        itv = qp.BootstrapProcessInterval(self.tmg, n_points=n_points, method="lifp", cptp=True, channel=self.chan)
        itv.setup(seed=seed)
        return itv

    def e2e_h2d(self):
        S = 4**self.cfg["n_qubits"]
        return S * self.povm.shape[-1] * 8 + 2 * self.centre.size * 8  # output-state Bloch vectors + centre Choi matrix

    api = "quantpy_b200.BootstrapProcessInterval(...).setup() + cl_to_dist"

    def kernel_breakdown(self):
        gpu, B = self.gpu, self.B
        out = {"sampler": gpu.time_alone(lambda: self.tmg.sample_counts(B, self.n_meas, self.povm, seed=1, device=True))}
        counts = self.tmg.sample_counts(B, self.n_meas, self.povm, seed=1, device=True)
        out["lifp_inversion"] = gpu.time_alone(lambda: self.tmg.point_estimate_batch(counts, cptp=False, device=True))
        out["lifp_inversion_and_cptp"] = gpu.time_alone(lambda: self.tmg.point_estimate_batch(counts, cptp=True, device=True))
        choi = self.tmg.point_estimate_batch(counts, cptp=True, device=True)
        out["distance"] = gpu.time_alone(lambda: gpu.engine.distance(choi, self.centre, "hs"))
        return out

    def roofline(self, kernels):
        gpu, cfg = self.gpu, self.cfg
        n = cfg["n_qubits"]
        s = 4**n
        cptp_ms = kernels["lifp_inversion_and_cptp"] - kernels["lifp_inversion"]
        its = self.mean_iterations() if self.last_iters is not None else None
        models = kernel_models().get("k_cptp", {})
        out = {"kernel": "k_cptp", "kernel_ms": cptp_ms, "mean_iterations": its, "bound": "fp64"}
        key = f"fp64_instructions_per_matrix_iteration_n{n}"
        if key in models and its:
            inst = models[key]
            peak = gpu.fp64_peak()
            ach = self.B * its * inst * 2.0 / (cptp_ms * 1e-3) / 1e12
            out.update(achieved=ach, peak=peak, unit="TFLOP/s (FP64 pipe slots x 2)", frac=ach / peak,
                       work_note=f"{inst} FP64 instructions per Choi matrix and alternating-projection iteration (ncu, "
                                 "profiles/kernel_models.json)", peak_source="qpb_fp64_fma_probe, measured in this run")
        else:
            out.update(achieved=None, peak=gpu.fp64_peak(), unit="TFLOP/s", frac=None,
                       work_note="no executed-instruction count committed for this kernel")
        g_ms = kernels["lifp_inversion"]
        S, K = s, self.povm.shape[0] * self.povm.shape[1]
        fl = 2.0 * self.B * (S * K) * (2 * s * s)
        out["dmma_inversion"] = {"kernel": "k_gemm_counts_dmma", "kernel_ms": g_ms, "bound": "fp64 tensor",
                                 "achieved": fl / (g_ms * 1e-3) / 1e12, "peak": gpu.fp64_peak(), "unit": "TFLOP/s",
                                 "frac": fl / (g_ms * 1e-3) / 1e12 / gpu.fp64_peak(),
                                 "work_note": "2 (S K) (2 d^4) flop per replica"}
        return out


def make_workload(gpu, cfg):
    return ProcessWorkload(gpu, cfg) if cfg["kind"] == "process" else StateWorkload(gpu, cfg)


def measure(gpu, cfg, steps, warmup, e2e_steps, with_cpu, cpu_seconds, with_roofline=True):
    """One config: device-resident throughput, end-to-end figure, dominant-kernel roofline, CPU baseline."""
    torch, lib, qpar = gpu.torch, gpu.lib, gpu.qpar
    world, rank = gpu.world, gpu.rank
    wl = make_workload(gpu, cfg)
    B = wl.B
    quiet_gc()   # before the warm-up, so that the device is busy right up to the first timed step
    for i in range(warmup):
        wl.step(i)
    # a config measured after a CPU-only phase (the previous config's cpu_baseline leg) finds the SM clock at idle:
    # keep the device busy for 50 ms more (untimed) so that the timed steps run at the clock the `clocks` key reports
    t_busy = time.perf_counter() + 0.05
    while time.perf_counter() < t_busy:
        wl.step(0)
        torch.cuda.synchronize()
    gpu.barrier()
    lib.qpb_reset_launch_count()
    evs = gpu.event_pairs(steps)
    gpu.barrier()
    for i, (e0, e1) in enumerate(evs):
        gpu.flush.zero_()                  # evict the previous step's data from the 126 MB L2 (not timed)
        e0.record()
        wl.step(warmup + i)
        e1.record()
    gpu.barrier()
    gc.enable()
    launches = int(lib.qpb_launch_count())
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    total_s = gpu.max_over_ranks(float(np.sum(step_ms)) / 1e3)
    value = world * B * steps / total_s
    mean_iters = wl.mean_iterations()

    # ---- e2e: the public API, host inputs -> host outputs, every step ------------------------------
    levels = np.linspace(1e-3, 1 - 1e-3, 1000)                          # ConfidenceInterval.__call__'s default levels
    traffic0 = dict(qpar.TRAFFIC)
    times, times_full = [], []
    warm_calls = 2  # two interval objects are alive at a time: both generations of buffers exist before timing
    quiet_gc()
    for i in range(warm_calls + e2e_steps):
        gpu.barrier()
        t0 = time.perf_counter()
        itv = wl.interval(B * world, seed=wl.seed + 100 + i)
        _ = itv.cl_to_dist(levels)       # the result a user reads: distances at the confidence levels, on the host
        # the call returned after a stream synchronisation with its result on this rank's host; its all-gather has
        # already tied the ranks together, and the reported time is the MAX over ranks of each rank's own sum
        t1 = time.perf_counter()
        _ = itv.dist                     # ... and all N sorted distances, as the reference's cl_to_dist owns them
        t2 = time.perf_counter()
        gpu.barrier()
        if i >= warm_calls:
            times.append(t1 - t0)
            times_full.append(t2 - t0)
    gc.enable()
    h2d = wl.e2e_h2d() + (qpar.TRAFFIC["h2d"] - traffic0["h2d"]) // (warm_calls + e2e_steps)
    d2h = (qpar.TRAFFIC["d2h"] - traffic0["d2h"]) // (warm_calls + e2e_steps) - B * world * 8
    e2e_s = gpu.max_over_ranks(float(np.sum(times)))
    e2e_full_s = gpu.max_over_ranks(float(np.sum(times_full)))
    e2e = {"value": world * B * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_call": 1e3 * e2e_s / e2e_steps,
           "ms_each": [round(1e3 * t, 3) for t in times], "api": wl.api,
           "note": "host in: Bloch vector, centre state, confidence levels; host out: the quantiles at the 1000 default levels; the sorted distances stay on the device",
           "dist_on_host": {"value": world * B * e2e_steps / e2e_full_s, "unit": UNIT,
                            "d2h_bytes_per_step": int(d2h + B * world * 8),
                            "note": "plus interval.dist: all N sorted distances on the host, as in the reference"}}

    rec = {"metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": 1e3 * total_s / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": workload_config(cfg, world), "gpu_launches": launches,
           "ms_each": [round(t, 4) for t in step_ms], "e2e": e2e}
    if mean_iters is not None:
        rec["mean_iterations"] = mean_iters
    if with_roofline and rank == 0:
        kernels = wl.kernel_breakdown()
        rec["kernels_ms_alone"] = kernels
        rec["roofline"] = wl.roofline(kernels)
    elif with_roofline:
        wl.kernel_breakdown()  # keep the ranks in lock step (no collectives inside, but equal GPU load)
    rec["cpu_baseline"] = cpu_baseline(cfg, cpu_seconds) if (with_cpu and rank == 0 and world == 1) else None
    return rec, wl


def strong_record(gpu, cfg, steps, warmup):
    """BASELINE configs[1] as written: the config's resamples GLOBALLY, sharded over the ranks; the timed region
    (CUDA events, max over ranks) holds sampler -> lin -> MLE -> distance on the shard, the sort of the shard, the
    all-gather and the merge of the sorted shards (interval.py:598-612 up to `dist.sort()`)."""
    torch, qpar = gpu.torch, gpu.qpar
    world, rank = gpu.world, gpu.rank
    n_total = cfg["resamples"]
    lo, hi = qpar.shard_bounds(n_total, rank, world)
    wl = StateWorkload(gpu, dict(cfg, resamples=hi - lo))

    def step(i):
        wl.plan.bootstrap_into(wl.bufs, wl.probs, wl.ref, wl.seed + i, lo, **wl.kw)
        return qpar.gather_sorted(wl.bufs["dist"], n_total)

    quiet_gc()
    full = None
    for i in range(max(warmup, 2)):
        # the previous step's result stays alive while the next one is produced, exactly as in the timed loop: both
        # generations of the gathered / merged buffers exist in torch's allocator before timing (the second timed step
        # of the 8-GPU record of round 2d paid a cudaMalloc: 1.43 ms against 0.83)
        full = step(i)
    gpu.barrier()
    evs = gpu.event_pairs(steps)
    for i, (e0, e1) in enumerate(evs):
        gpu.flush.zero_()
        gpu.barrier()       # ranks enter the step together, so the collective does not absorb launch skew
        e0.record()
        full = step(warmup + i)
        e1.record()
    gpu.barrier()
    gc.enable()
    each = [e0.elapsed_time(e1) for e0, e1 in evs]
    total_s = gpu.max_over_ranks(float(np.sum(each)) / 1e3)
    srt = full.cpu().numpy()
    return {"value": n_total * steps / total_s, "unit": UNIT, "scaling": "strong", "global_resamples": n_total,
            "resamples_per_gpu": hi - lo, "n_gpus": world, "steps": steps, "ms_per_step": 1e3 * total_s / steps,
            "ms_each": [round(t, 4) for t in each], "ms_per_step_median": float(np.median(each)),
            "timed_region": "sampler + lin + MLE + distance on the shard, shard sort, one all-gather, merge of the sorted "
                            "shards; CUDA events, max over ranks, L2 flushed and ranks aligned before every step",
            "sorted_ok": bool(np.all(np.diff(srt) >= 0) and len(srt) == n_total)}


def run_ours(args, rank, local_rank, world):
    gpu = Gpu(rank, local_rank, world)
    cfg = resolve_config(args)
    clocks = clock_sampler(local_rank)
    # warm the device up before the clock sampler starts, so the samples are under load
    StateWorkload(gpu, dict(resolve_config(args, "c2"), resamples=20000)).step(0)
    gpu.barrier()
    if rank == 0:
        clocks.start()
    rec, _ = measure(gpu, cfg, args.steps, args.warmup, e2e_steps=args.steps, with_cpu=not args.no_cpu_baseline,
                     cpu_seconds=args.cpu_seconds)
    clock_info = clocks.stop() if rank == 0 else None
    rec["clocks"] = clock_info
    if cfg["kind"] == "state" and cfg["method"] == "mle":
        rec["strong"] = strong_record(gpu, cfg, args.steps, args.warmup)
    if world == 1 and args.config == "c2" and not args.no_configs:
        others = []
        for name in ("c1", "c3", "c4", "c5-1q", "c5-2q"):
            sub = resolve_config(args, name)
            heavy = name in ("c4",)
            r, _ = measure(gpu, sub, steps=3 if heavy else 10, warmup=3, e2e_steps=2 if heavy else 5,
                           with_cpu=not args.no_cpu_baseline, cpu_seconds=4.0)
            r["name"] = name
            others.append(r)
        rec["configs"] = others
    if rank == 0:
        print(json.dumps(rec), flush=True)
    if world > 1:
        gpu.dist.barrier()
        gpu.dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
