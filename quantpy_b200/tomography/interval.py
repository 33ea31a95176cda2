"""Bootstrap confidence intervals -- drop-in for the bootstrap functors of
quantpy/tomography/interval.py (ConfidenceInterval :19-56, BootstrapStateInterval :542-612,
BootstrapProcessInterval :615-685).

The reference runs `n_points` serial Python iterations of experiment -> point_estimate -> dst.  Here
the whole loop is one batched pass on the GPU (engine.StatePlan.bootstrap / ProcessPlan.lifp), sharded
over the ranks of torch.distributed when a process group exists, followed by one all-gather of the
per-sample distances, a sort and the same interp1d quantile function.
"""

from abc import ABC, abstractmethod
from enum import Enum, auto

import numpy as np

from .. import _native as nt
from .. import engine, parallel
from .state import dst_kind


class Mode(Enum):
    STATE = auto()
    CHANNEL = auto()


def _pop_hidden_keys(kwargs):
    """Constructor kwargs become attributes, minus self/tmg/dunder names (interval.py:858-865)."""
    return {k: v for k, v in kwargs.items() if k not in ("self", "tmg") and not k.startswith("__")}


_DEFAULT_LEVELS = np.linspace(1e-3, 1 - 1e-3, 1000)  # ConfidenceInterval.__call__'s default (interval.py:41-42)
_DEFAULT_LEVELS.setflags(write=False)
_NO_LEVELS = np.zeros(0)


class ConfidenceInterval(ABC):
    """Functor: interval(conf_levels) -> (distances, conf_levels); `setup` runs lazily on first call."""

    EPS = 1e-15
    _setup_takes_levels = False

    def __init__(self, tmg, **kwargs):
        self.tmg = tmg
        if hasattr(tmg, "state"):
            self.mode = Mode.STATE
        elif hasattr(tmg, "channel"):
            self.mode = Mode.CHANNEL
        else:
            raise ValueError()
        for name, value in kwargs.items():
            setattr(self, name, value)

    def __call__(self, conf_levels=None):
        if conf_levels is None:
            conf_levels = _DEFAULT_LEVELS.copy()
        if not hasattr(self, "cl_to_dist"):
            if self._setup_takes_levels:
                self.setup(conf_levels=conf_levels)  # the levels travel with the set-up call
            else:
                self.setup()
        return self.cl_to_dist(conf_levels), conf_levels

    @abstractmethod
    def setup(self):
        """Configure the confidence interval."""

    def _finish(self, local_dist, n_points):
        """All-gather the per-rank distances, keep them (sorted) and build the quantile function."""
        # sorted on the device and left there: the quantile function fetches what a call needs, `dist` the rest
        self.cl_to_dist = parallel.quantile_function(parallel.gather_sorted(local_dist, n_points), presorted=True)

    @property
    def dist(self):
        """The sorted distances behind `cl_to_dist` (host array; copied from the device on first use)."""
        return self.cl_to_dist.y


class MomentInterval(ConfidenceInterval):
    def __init__(self, tmg, distr_type="gamma"):
        """Analytic interval from the first two moments of the squared estimation error
        (quantpy/tomography/interval.py:59-110).  distr_type: 'gamma' | 'norm' | 'exp'."""
        super().__init__(tmg, **_pop_hidden_keys(locals()))

    def setup(self):
        import scipy.stats as sts

        from ..geometry import hs_dst, trace_dst
        from ..routines import _left_inv

        if self.mode == Mode.STATE:
            dim = 2**self.tmg.state.n_qubits
            n_measurements = self.tmg.n_measurements
            frequencies = self.tmg.results / n_measurements[:, None]
            povm = np.asarray(self.tmg.povm_matrix, dtype=np.float64)
            inv = _left_inv(povm.reshape(-1, povm.shape[-1])) / dim
        else:
            dim = 4**self.tmg.channel.n_qubits
            first = self.tmg.tomographs[0]
            n_measurements = first.n_measurements
            frequencies = np.vstack([t.results / n_measurements[:, None] for t in self.tmg.tomographs])
            povm = np.asarray(first.povm_matrix, dtype=np.float64)
            flat = povm.reshape(-1, povm.shape[-1])
            states = np.asarray([rho.T.bloch for rho in self.tmg.input_basis.elements])
            chan = np.einsum("sd,pi->spdi", states, flat).reshape(len(states) * len(flat), -1)
            inv = _left_inv(chan) / dim
        inv = inv.reshape(inv.shape[0], frequencies.shape[0], frequencies.shape[1])
        weights = np.einsum("aij,akl->ijkl", inv, inv)  # one-time set-up on the host
        self.mean, self.variance = engine.l2_moments(frequencies, n_measurements[0], weights)  # qpb_l2_moments
        if self.distr_type == "norm":
            distr = sts.norm(loc=self.mean, scale=np.sqrt(self.variance))
        elif self.distr_type == "gamma":
            scale = self.variance / self.mean
            distr = sts.gamma(a=self.mean / scale, scale=scale)
        elif self.distr_type == "exp":
            distr = sts.expon(scale=self.mean)
        else:
            raise NotImplementedError(f"Unsupported distribution type {self.distr_type}")
        if self.tmg.dst == hs_dst:
            alpha = np.sqrt(dim / 2)
        elif self.tmg.dst == trace_dst:
            alpha = dim / 2
        else:
            raise NotImplementedError()
        self.cl_to_dist = lambda cl: np.sqrt(distr.ppf(cl)) * alpha


class BootstrapStateInterval(ConfidenceInterval):
    def __init__(self, tmg, n_points=1000, method="lin", physical=True, init="lin", tol=1e-3, max_iter=100,
                 state=None):
        """Parametric bootstrap around `state` (default: the tomograph's reconstructed state) with the
        tomograph's POVM and shot numbers; kwargs as StateTomograph.point_estimate."""
        super().__init__(tmg, **_pop_hidden_keys(locals()))

    _setup_takes_levels = True

    def setup(self, seed=None, conf_levels=None):
        """interval.py:583-612.  `conf_levels` (optional; default: the levels `__call__` uses) are evaluated by the
        same device call that runs the bootstrap, so a following `cl_to_dist(conf_levels)` costs nothing."""
        if self.mode == Mode.CHANNEL:
            raise NotImplementedError("This interval works only for state tomography")
        if self.state is None:
            if hasattr(self.tmg, "reconstructed_state"):
                self.state = self.tmg.reconstructed_state
            else:
                self.state = self.tmg.point_estimate(method=self.method, physical=self.physical, init=self.init,
                                                     tol=self.tol, max_iter=self.max_iter)
        if self.method not in ("lin", "mle", "mle-constr"):
            raise ValueError("Invalid value for argument `method`")
        method = "mle" if self.method == "mle-constr" else self.method  # same maximiser, same kernel
        rank, size = parallel.world()
        lo, hi = parallel.shard_bounds(self.n_points, rank, size)
        # a key drawn from this rank's np.random stream must be agreed on; an explicit seed is the caller's (SPMD)
        seed = parallel.broadcast_seed(engine.next_seed()) if seed is None else int(seed)
        plan = engine.state_plan(self.tmg.povm_matrix, self.tmg.n_measurements)
        kind = dst_kind(self.tmg.dst)
        if kind is not None and hi > lo and engine.FUSED_INTERVAL:
            # built-in distance: ONE library call with host inputs and outputs (qpb_bootstrap_state_interval) runs this
            # rank's shard -- upload, probabilities, bootstrap, shard sort and, on one GPU, the quantiles as well
            if size > 1:
                levels = _NO_LEVELS    # several GPUs: the quantiles follow the gather
            elif conf_levels is None:
                levels = _DEFAULT_LEVELS
            else:
                levels = np.asarray(conf_levels, dtype=np.float64)
                if levels.size and (levels.min() < 0 or levels.max() > 1):
                    levels = _NO_LEVELS  # cl_to_dist raises interp1d's error when it is called
            q, dist, self._iters_dev = engine.bootstrap_interval(
                plan, self.state.bloch, self.state.matrix, hi - lo, seed, lo, method, self.physical, self.init,
                self.max_iter, self.tol, kind, levels)
            parallel.TRAFFIC["h2d"] += (3 * plan.D + levels.size) * 8
            parallel.TRAFFIC["d2h"] += levels.size * 8
            if size > 1:
                dist = parallel.gather_sorted(dist, self.n_points, presorted=True)
            self.cl_to_dist = parallel.quantile_function(dist, presorted=True)
            if size == 1:
                self.cl_to_dist.prime(levels, q.reshape(levels.shape))
            return
        probs = plan.probabilities(self.state.bloch)[0]
        out = plan.bootstrap(probs, hi - lo, seed, lo, self.state.matrix, method, self.physical, self.init,
                             self.max_iter, self.tol, kind or "hs", keep=kind is None)
        if kind is None:  # user-supplied measure: evaluate it on the reconstructed batch
            from ..qobj import Qobj

            torch = nt.torch_cuda()
            rhos = nt.complex_to_host(out["rho"])
            local = torch.tensor([float(self.tmg.dst(Qobj(r), self.state)) for r in rhos], dtype=torch.float64,
                                 device="cuda")
        else:
            local = out["dist"]
        self._iters_dev = out["iters"]  # copied to the host only if someone asks (see `iters`)
        self._finish(local, self.n_points)

    @property
    def iters(self):
        """R.rho.R iteration count of every local bootstrap sample (zeros for 'lin')."""
        return self._iters_dev.cpu().numpy()


class MHMCStateInterval(ConfidenceInterval):
    def __init__(self, tmg, n_points=1000, step=0.01, burn_steps=1000, thinning=1, warm_start=False,
                 use_new_estimate=False, state=None, verbose=False):
        """Metropolis-Hastings samples from the likelihood, distances to the point estimate
        (quantpy/tomography/interval.py:689-759, chain: quantpy/mhmc.py:48-119).

        The chain is the reference's: packed Cholesky vector on the unit sphere, proposal
        (x + step*delta)/|x + step*delta| with delta ~ N(0, I), target exp(-tmg._nll(x)).  Its noise is drawn on
        the host from the legacy np.random stream in the reference's order (burn-in deltas, burn-in uniforms,
        sampling deltas, sampling uniforms), so `np.random.seed(s)` reproduces the reference's chain; the
        steps themselves run in one CUDA kernel (qpb_mhmc_state).  `verbose` is accepted and ignored."""
        super().__init__(tmg, **_pop_hidden_keys(locals()))

    def setup(self):
        if self.mode == Mode.CHANNEL:
            raise NotImplementedError("This interval works only for state tomography")
        if not self.use_new_estimate:
            self.state = self.tmg.reconstructed_state
        elif self.state is None:
            self.state = self.tmg.point_estimate(method="mle", physical=True)
        plan = engine.state_plan(self.tmg.povm_matrix, self.tmg.n_measurements)
        dim = plan.D
        warm = self.warm_start and hasattr(self, "_x_t")
        x_init = self._x_t if warm else engine.cholesky_vector(self.state.matrix)
        burn = 0 if warm else int(self.burn_steps)
        total = int(self.n_points) * int(self.thinning)
        # mhmc.py:70-71, 89-90: jump_distr.rvs (SciPy's frozen N(0, I) calls multivariate_normal of the global
        # legacy stream) and np.random.rand, burn-in first
        mean, cov = np.zeros(dim), np.eye(dim)
        parts_d, parts_u = [], []
        for size in ([burn] if burn else []) + [total]:
            parts_d.append(np.random.multivariate_normal(mean, cov, size).reshape(size, dim))
            parts_u.append(np.random.rand(size))
        out = engine.mhmc_chains(plan, self.tmg.results, x_init[None], self.n_points, self.step, burn, self.thinning,
                                 np.concatenate(parts_d)[None], np.concatenate(parts_u)[None])
        self._x_t = out["x_final"][0].cpu().numpy()
        self.acceptance_rate = float(out["accepted"][0].item()) / total
        kind = dst_kind(self.tmg.dst)
        samples = out["samples"][0]
        if kind is None:
            from ..qobj import Qobj

            mats = nt.complex_to_host(samples)
            dist = np.sort([float(self.tmg.dst(Qobj(m), self.state)) for m in mats])
            self.cl_to_dist = parallel.quantile_function(dist, presorted=True)
            return
        self.cl_to_dist = parallel.quantile_function(engine.distance(samples, self.state.matrix, kind))


class BootstrapProcessInterval(ConfidenceInterval):
    def __init__(self, tmg, n_points=1000, method="lifp", cptp=True, tol=1e-10, channel=None,
                 states_est_method="lin", states_physical=True, states_init="lin"):
        """Parametric bootstrap around `channel` (default: the tomograph's reconstructed channel);
        kwargs as ProcessTomograph.point_estimate.  Supported for n_qubits <= 2 (see ProcessTomograph.sample_counts).
        Deviation from the reference: with method='states' the replicas use `states_est_method` (the reference's
        loop never forwards it, quantpy/tomography/interval.py:676-681, so it always bootstraps with 'lin')."""
        super().__init__(tmg, **_pop_hidden_keys(locals()))

    def setup(self, seed=None):
        if self.mode == Mode.STATE:
            raise NotImplementedError("This interval works only for process tomography")
        if self.channel is None:
            if hasattr(self.tmg, "reconstructed_channel"):
                self.channel = self.tmg.reconstructed_channel
            else:
                self.channel = self.tmg.point_estimate(method=self.method, states_physical=self.states_physical,
                                                       states_init=self.states_init, cptp=self.cptp)
        if self.method not in ("lifp", "states"):
            raise ValueError("Incorrect value for argument `method`")
        rank, size = parallel.world()
        lo, hi = parallel.shard_bounds(self.n_points, rank, size)
        seed = parallel.broadcast_seed(engine.next_seed()) if seed is None else int(seed)
        first = self.tmg.tomographs[0]
        # same input basis, POVM tables and shot bookkeeping, another channel: no sampling, no RNG draw, no
        # re-derivation of the basis decompositions (the batch estimators read only POVM and shots from `tomographs`)
        boot = self.tmg._with_channel(self.channel)
        centre = self.channel.choi.matrix
        kind = dst_kind(self.tmg.dst)
        torch = nt.torch_cuda()
        counts = boot.sample_counts(hi - lo, first.n_measurements, first.povm_matrix, seed, lo, device=True)
        if self.method == "lifp":
            choi = boot.point_estimate_batch(counts, cptp=self.cptp, device=True)
        else:
            # the reference's bootstrap forwards only states_physical / states_init (interval.py:676-681), so its
            # replicas always use the default states_est_method='lin'; here the constructor's value is honoured
            choi = boot.point_estimate_states_batch(counts, cptp=self.cptp, method=self.states_est_method,
                                                    physical=self.states_physical, init=self.states_init, device=True)
        if kind is not None:
            local = engine.distance(choi, centre, kind)
        else:
            from ..qobj import Qobj

            local = torch.tensor([float(self.tmg.dst(Qobj(c), self.channel.choi))
                                  for c in nt.complex_to_host(choi)], dtype=torch.float64, device="cuda")
        self._finish(local, self.n_points)


def _out_of_scope(name):
    class _Unavailable(ConfidenceInterval):
        def __init__(self, *args, **kwargs):
            raise NotImplementedError(
                f"{name} is outside the B200 bootstrap hot path (analytic / LP / MCMC interval, see DESIGN.md); "
                "use the reference implementation for it."
            )

        def setup(self):  # pragma: no cover
            raise NotImplementedError

    _Unavailable.__name__ = _Unavailable.__qualname__ = name
    return _Unavailable


MomentFidelityStateInterval = _out_of_scope("MomentFidelityStateInterval")
MomentFidelityProcessInterval = _out_of_scope("MomentFidelityProcessInterval")
SugiyamaInterval = _out_of_scope("SugiyamaInterval")
PolytopeStateInterval = _out_of_scope("PolytopeStateInterval")
PolytopeProcessInterval = _out_of_scope("PolytopeProcessInterval")
HolderInterval = _out_of_scope("HolderInterval")
MHMCProcessInterval = _out_of_scope("MHMCProcessInterval")
