import numpy as np
z = np.load("/tmp/study/traj.npz"); n = z["iters"]; lam = z["lam_min"]; B = len(n)
# queue order: small positive first ... large positive, then negatives
key = np.where(lam > 0, lam, np.inf)
order = np.argsort(key, kind="stable")
rank = np.empty(B, int); rank[order] = np.arange(B)
print("positive fraction", (lam > 0).mean())
for N in (300, 400, 450, 500, 550, 600, 674, 800):
    long_ = n > N
    r = np.sort(rank[long_]) / B
    if len(r) == 0: continue
    print(f"n > {N}: {long_.sum():5d} samples; queue-position quantiles 50% {np.percentile(r,50):.3f} 90% {np.percentile(r,90):.3f} 99% {np.percentile(r,99):.3f} max {r.max():.3f}")
for q in (0.01, 0.02, 0.03, 0.05, 0.1, 0.2):
    top = rank < q * B
    print(f"top {q:.2f} of the queue: mean its {n[top].mean():.0f}, P(n>300) {np.mean(n[top]>300):.2f}, P(n>450) {np.mean(n[top]>450):.2f}; outside: max its {n[~top].max()}, P(n>450) {np.mean(n[~top]>450):.4f}")
