import numpy as np
z = np.load("/tmp/study/traj.npz"); n = z["iters"]; d = z["dels"]; B = len(n)
tol2 = 1e-12
TS, TW = 2.6, 0.66
def report(name, A):   # A = hand-over iteration per sample (>= n: never)
    A = np.minimum(A, n)
    t = A * TS + (n - A) * TW
    print(f"{name:60s} handed {np.mean(A<n):.4f}  W its/sample {np.mean(n-A):6.2f}  chain max {t.max():7.1f} us  p99.9 {np.percentile(t,99.9):7.1f}")
for age in (250, 300, 350, 400, 450, 500):
    report(f"age {age}", np.full(B, age))
# predictor: first checkpoint c >= cmin where it + k*rem > N
with np.errstate(all="ignore"):
    r = d[:, 1:] / d[:, :-1]
    rem = np.where(r < 1, 32 * np.log(tol2 / d[:, 1:]) / np.log(r), 1e9)   # index j -> checkpoint j+1
for cmin in (4, 5, 6, 7, 8):
    for N in (350, 400, 450, 500):
        for k in (1.0, 1.15, 1.3):
            for cap in (400, 500):
                A = np.full(B, cap)
                for c in range(cmin, 31):
                    it = 32 * c
                    if it >= cap: break
                    pred = it + k * rem[:, c - 1]
                    hit = (n > it) & (pred > N) & (A == cap)
                    A[hit] = it
                report(f"predict from {32*cmin}: it + {k}*rem > {N}, fallback age {cap}", A)
