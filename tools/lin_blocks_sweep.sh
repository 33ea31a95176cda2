# k_lin_project_small<2>: resident blocks per SM (register cap) against the time of lin + start order at 1e5 samples.
# Rebuilds lin_small.o on the GPU box for every setting (the image has nvcc).
cd $GRAFT_REPO_ROOT
cat > /tmp/time_lin.py <<'PY'
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import quantpy_b200 as qp
from quantpy_b200 import engine
rng = np.random.default_rng(0)
g = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)); rho = g @ g.conj().T; rho /= np.trace(rho)
povm = qp.generate_measurement_matrix("proj", 2)
plan = engine.state_plan(povm, np.ones(1) * 10000)
probs = plan.probabilities(qp.Qobj(rho).bloch)[0]
for B in (100000, 12500):
    c = plan.sample(probs, B, 1, 0)
    for rep in range(3): r, o = plan.lin_ordered(c)
    torch.cuda.synchronize()
    ms = []
    for rep in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r, o = plan.lin_ordered(c); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    print(f"  B={B}: lin + order min {min(ms):.4f} ms median {np.median(ms):.4f} checksum {r.sum().item():.12f}")
PY
for mb in 2 3 4 5 6; do
  (cd quantpy_b200/csrc && make -B lin_small.o EXTRA="-DLIN_MIN_BLOCKS=$mb -Xptxas -v" 2>&1 | grep -A2 "k_lin_project_smallILi2" | grep -E "registers|spill" | tr '\n' ' ' && make > /dev/null 2>&1)
  echo "min blocks $mb"; python /tmp/time_lin.py
done
(cd quantpy_b200/csrc && make -B lin_small.o > /dev/null 2>&1 && make > /dev/null 2>&1)
