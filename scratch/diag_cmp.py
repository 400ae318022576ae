import numpy as np, sys
a, b = sys.argv[1], sys.argv[2]
for niter in (11, 12, 14, 17, 20, 25, 30, 40):
    x = np.load(f"/tmp/diag_{a}_{niter}.npy"); y = np.load(f"/tmp/diag_{b}_{niter}.npy")
    d = np.abs(x - y).max(axis=2)
    bad = np.argwhere(d > 0)
    print(niter, "max", float(d.max()), "ndiff", len(bad), "first", bad[:8].tolist(), "rows", sorted(set(bad[:, 0].tolist()))[:12])
    if len(bad):
        j, i = bad[0]; print("    ", a, x[j, i], b, y[j, i])
