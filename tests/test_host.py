"""CPU tests: host-side mirror of the reference API, the C-ABI library's exports, and the
multi-rank plumbing (gloo, world_size 2).  No kernel is launched here."""

import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import quantpy_b200 as qp
from quantpy_b200 import _native, parallel
from quantpy_b200.routines import _left_inv, _mat2vec, _out_ptrace_oper, _vec2mat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_qobj_representations(golden):
    g = golden("api")
    for n in (1, 2, 3):
        q = qp.Qobj(g[f"rho_{n}"])
        assert q.n_qubits == n
        assert np.allclose(q.bloch, g[f"bloch_{n}"], atol=1e-15)
        assert np.allclose(qp.Qobj(g[f"bloch_{n}"]).matrix, g[f"back_{n}"], atol=1e-15)
    assert np.allclose(qp.Qobj([0, 1], is_ket=True).bloch, g["ket01_bloch"])
    assert np.allclose(qp.Qobj([0.5, 0, 0, 0.5]).matrix, [[1, 0], [0, 0]])
    short = qp.Qobj([0.1, 0.2, 0.3])  # 4^n - 1 entries: identity coefficient is prepended
    assert np.allclose(short.bloch, [0.5, 0.1, 0.2, 0.3])
    rho2 = qp.Qobj(g["rho_2"])
    assert np.allclose(rho2.ptrace([0]).matrix, g["ptrace_keep0"])
    assert np.allclose(rho2.ptrace([1]).matrix, g["ptrace_keep1"])
    assert rho2.is_density_matrix() and not rho2.is_pure()
    assert qp.Qobj([1, 0], is_ket=True).is_pure()
    with pytest.raises(ValueError):
        qp.Qobj(np.zeros((2, 2, 2)))


def test_qobj_arithmetic():
    a = qp.Qobj(np.array([[1, 0], [0, 0]], dtype=complex))
    b = qp.Qobj(np.array([[0, 0], [0, 1]], dtype=complex))
    assert np.allclose((a + b).matrix, np.eye(2))
    assert np.allclose((0.5 * a + b * 0.5).matrix, np.eye(2) / 2)
    assert np.allclose((np.complex128(2) * a).matrix, 2 * a.matrix)
    assert np.allclose((a / a.trace()).matrix, a.matrix)
    assert np.allclose(a.kron(b).matrix, np.kron(a.matrix, b.matrix))
    assert np.allclose(qp.kron(a, b).matrix, np.kron(a.matrix, b.matrix))
    c = a.copy()
    c += b
    assert np.allclose(c.matrix, np.eye(2)) and np.allclose(a.matrix, [[1, 0], [0, 0]])
    assert a == qp.Qobj(a) and a != b
    with pytest.raises(ValueError):
        a * "x"


def test_measurement_matrices(golden):
    g = golden("api")
    for name in ("proj", "proj-set", "proj4", "sic"):
        for n in (1, 2):
            assert np.array_equal(qp.generate_measurement_matrix(name, n), g[f"povm_{name}_{n}"])
    one = qp.generate_measurement_matrix("proj", 1)[0]
    assert np.array_equal(qp.generate_measurement_matrix(one, 2), g["povm_proj_2"])
    full = g["povm_proj-set_2"]
    assert qp.generate_measurement_matrix(full, 2) is full
    assert qp.generate_measurement_matrix(full[0], 2).shape == (1, 4, 16)
    for bad in ("nope", np.zeros((3, 5)), 3):
        with pytest.raises(ValueError):
            qp.generate_measurement_matrix(bad, 2)


def test_channel_and_basis(golden):
    g = golden("api")
    chan = qp.channel.depolarizing(p=0.3, n_qubits=1)
    assert np.allclose(chan.choi.matrix, g["dep03_choi"], atol=1e-15)
    by_choi = qp.Channel(g["dep03_choi"])
    assert np.allclose(by_choi.transform(qp.Qobj(g["rho_1"])).matrix, g["dep03_apply"], atol=1e-15)
    assert np.allclose(qp.operator.Z.as_channel().choi.matrix, g["zchan_choi"])
    assert chan.is_cptp() and not qp.Channel(-g["dep03_choi"]).is_cptp(verbose=False)
    kraus = qp.Channel(chan.kraus)
    assert np.allclose(kraus.choi.matrix, g["dep03_choi"], atol=1e-12)
    sic = [qp.Qobj(b) for b in np.squeeze(qp.generate_measurement_matrix("sic", 1))]
    basis = qp.basis.Basis(sic)
    assert np.allclose(basis.gram, g["basis_gram"], atol=1e-15)
    coefs = basis.decompose(qp.Qobj(g["rho_1"]))
    assert np.allclose(coefs, g["basis_decomp"], atol=1e-13)
    with pytest.raises(ValueError):
        qp.Channel(lambda r: r)
    with pytest.raises(ValueError):
        qp.ProcessTomograph(chan, input_states=sic[:3])


def test_distances_host(golden):
    g = golden("api")
    a, b = g["dst_a"], g["dst_b"]
    got = [qp.hs_dst(a, b), qp.trace_dst(qp.Qobj(a), qp.Qobj(b)), qp.if_dst(a, b)]
    assert np.allclose(got, g["dst_vals"], atol=1e-13)
    assert qp.hs_dst(a, a) == 0 and isinstance(qp.hs_dst(a, a), int)
    assert np.isclose(qp.product(a, b), np.trace(a @ b.conj().T))


def test_routines_match_oracle():
    from oracle import process as oproc

    rng = np.random.default_rng(0)
    m = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    assert np.array_equal(_mat2vec(m), oproc.mat2vec(m)) and np.array_equal(_vec2mat(_mat2vec(m)), m)
    for n in (1, 2):
        assert np.array_equal(_out_ptrace_oper(n), oproc.ptrace_operator(n))
    a = rng.normal(size=(6, 4)) + 1j * rng.normal(size=(6, 4))
    assert np.allclose(_left_inv(a) @ a, np.eye(4))
    assert np.allclose(qp.generate_pauli(2)[7], np.kron(qp.generate_pauli(1)[1], qp.generate_pauli(1)[3]))


def test_out_of_scope_intervals_say_so():
    class Fake:
        state = None

    for cls in (qp.MomentFidelityStateInterval, qp.SugiyamaInterval, qp.MHMCProcessInterval, qp.HolderInterval):
        with pytest.raises(NotImplementedError):
            cls(Fake())


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, never compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    tmg = qp.StateTomograph(qp.Qobj([0.5, 0, 0, 0.5]))
    with pytest.raises(qp.NativeError):
        tmg.experiment(100)
    tmg.povm_matrix = qp.generate_measurement_matrix("proj-set", 1)
    tmg.results = np.array([[50, 50], [50, 50], [100, 0]])
    with pytest.raises(qp.NativeError):
        tmg.point_estimate("lin")
    with pytest.raises(qp.NativeError):
        qp.BootstrapStateInterval(tmg, n_points=4, state=tmg.state).setup()


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "quantpy_b200")):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, name)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), name
                assert "/root/reference" not in text, name


# ----------------------------------------------------------------------------- C ABI

def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "quantpy_b200.h")).read()
    declared = set(re.findall(r"QPB_API\s+[\w\s\*]+?\b(qpb_\w+)\s*\(", header))
    assert len(declared) >= 26
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    lib = _native.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.qpb_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (qpb_\w+)", out))
    assert exported == declared
    lib.qpb_reset_launch_count()
    assert lib.qpb_launch_count() == 0
    # argument validation happens before any CUDA call: usable without a GPU
    assert lib.qpb_distance(3, 1, None, None, 0, None, None) < 0
    assert b"unsupported" in lib.qpb_last_error()
    handle = ctypes.c_void_p()
    assert lib.qpb_state_plan_create(ctypes.byref(handle), 7, 4, None, None, None) < 0


# ----------------------------------------------------------------------------- multi-rank plumbing

def test_shard_bounds_partition():
    for n, w in ((100000, 8), (10, 3), (5, 8), (0, 2)):
        cuts = [parallel.shard_bounds(n, r, w) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        sizes = [hi - lo for lo, hi in cuts]
        assert max(sizes) - min(sizes) <= 1


WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from quantpy_b200 import parallel
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, size = parallel.world()
n = 11
lo, hi = parallel.shard_bounds(n, rank, size)
seed = parallel.broadcast_seed(1234 + rank)
assert seed == 1234
# each rank "computes" the statistic of its slice of global sample indices
local = torch.arange(lo, hi, dtype=torch.float64) * 0.5 + seed
full = parallel.all_gather_concat(local, n)
assert torch.equal(full, torch.arange(n, dtype=torch.float64) * 0.5 + 1234), full
q = parallel.quantile_function(full.numpy()[::-1])
assert np.isclose(q(0.0), 1234.0) and np.isclose(q(1.0), 1239.0) and np.isclose(q(0.5), 1236.5)
# the interval's own path: every rank contributes an unsorted shard, all get the globally sorted vector
mixed = torch.tensor([((7 * i) % 11) * 0.25 for i in range(lo, hi)], dtype=torch.float64)
merged = parallel.gather_sorted(mixed, n)
assert torch.equal(merged, torch.sort(torch.tensor([((7 * i) % 11) * 0.25 for i in range(n)], dtype=torch.float64)).values)
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_all_gather_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o


def test_frequency_division_formula_is_correctly_rounded():
    """csrc/common.cuh::FreqDiv turns counts/total into one reciprocal plus q = c*r; q + fma(-q, total, c)*r.
    Emulated in exact rational arithmetic, the result has the bits of the IEEE quotient the reference computes
    (state.py:193, 227) for every count of realistic shot totals."""
    from fractions import Fraction as Fr

    rng = np.random.default_rng(0)
    totals = [3, 7, 1000, 9999, 10000, 30000, 90000, 12960000] + [int(t) for t in rng.integers(1, 10**7, 40)]
    for t in totals:
        inv = float(Fr(1) / t)
        counts = range(t + 1) if t <= 1000 else [int(c) for c in rng.integers(0, t + 1, 400)]
        for c in counts:
            q = float(Fr(c) * Fr(inv))
            res = float(Fr(c) - Fr(q) * t)           # fma(-q, total, c): exact product, one rounding
            got = float(Fr(q) + Fr(res) * Fr(inv))   # fma(res, inv, q)
            assert got == c / t, (c, t)


def test_quantile_function_equals_scipy_interp1d():
    """parallel.QuantileFunction is the object the reference builds with interp1d(linspace(0, 1, N), sorted dist)
    (interval.py:610-612): same values on random grids, same `.x` / `.y`, same ValueError outside [0, 1]."""
    from scipy.interpolate import interp1d

    from quantpy_b200 import parallel

    rng = np.random.default_rng(3)
    for n in (2, 3, 17, 1000):
        dist = np.sort(rng.random(n))
        ref = interp1d(np.linspace(0, 1, n), dist)
        q = parallel.quantile_function(dist[::-1].copy())
        levels = np.concatenate([[0.0, 1.0], rng.random(200), np.linspace(0, 1, n)])
        assert np.allclose(q(levels), ref(levels), rtol=0, atol=1e-15)
        assert np.array_equal(q.x, ref.x) and np.array_equal(q.y, ref.y)
        assert q(np.array([[0.25, 0.5]])).shape == (1, 2)
        for bad in (-1e-9, 1 + 1e-9):
            with pytest.raises(ValueError):
                q(bad)
            with pytest.raises(ValueError):
                ref(bad)


def test_quantile_function_answers_the_set_up_levels_from_the_primed_result(monkeypatch):
    """Device-resident sorted values: a call with exactly the levels of the set-up call (same bytes, same shape) is
    answered from what that call brought back; anything else goes to qpb_quantiles_host.  Range errors come first and
    read like interp1d's; empty and scalar level arrays pass through."""
    from quantpy_b200 import engine, parallel

    class FakeDeviceArray:
        is_cuda = True

        def numel(self):
            return 1000

    asked = []
    monkeypatch.setattr(engine, "quantiles_host", lambda dev, levels: asked.append(levels.copy()) or levels * 2.0)
    q = parallel.QuantileFunction(FakeDeviceArray())
    levels = np.linspace(1e-3, 1 - 1e-3, 1000)
    values = np.arange(1000.0)
    q.prime(levels, values)
    out = q(levels.copy())
    assert np.array_equal(out, values) and not asked
    out[0] = -1.0                                   # the caller owns what it gets
    assert q(levels)[0] == 0.0
    assert np.array_equal(q(list(levels)), values) and not asked          # any array-like with the same content
    assert np.array_equal(q(levels.reshape(10, 100)), levels.reshape(10, 100) * 2.0) and len(asked) == 1  # other shape
    other = levels.copy()
    other[500] = np.nextafter(other[500], 1.0)
    assert np.array_equal(q(other), other * 2.0) and len(asked) == 2
    assert q(np.zeros(0)).shape == (0,) and q(0.5) == 1.0
    with pytest.raises(ValueError, match="below the interpolation range"):
        q([0.5, -1e-9, 1.5])
    with pytest.raises(ValueError, match="above the interpolation range"):
        q([0.5, 1 + 1e-9])
    assert len(asked) == 4


def test_shard_bounds_partition_any_range():
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from quantpy_b200 import parallel

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 10**6), st.integers(1, 64))
    def check(n, world):
        edges = [parallel.shard_bounds(n, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
        sizes = [hi - lo for lo, hi in edges]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)

    check()
