"""Numerical check: lexicographic SOR sweep == overlapped local sweeps (halo W/S/N) up to rounding."""
import sys, numpy as np
sys.path.insert(0, '.')
from oracle import refapi
from opticalflow2d_b200 import synthetic as S
bits = int(sys.argv[1]) if len(sys.argv) > 1 else 32
o = refapi.get("oracle", bits)
nx, ny = 200, 160
R, T = S.make_pair(nx, ny, "lattice", shift=(1.5, -0.75))
R = o.set_image(R); T = o.set_image(T)
params = [float(x) for x in sys.argv[2:5]] if len(sys.argv) > 4 else [1.0, 0.25, 0.66]
u0 = np.zeros((ny, nx, 2), o.real)
u_old = o.solver_steps(2, params, R, T, u0, 5)
u_new = o.solver_steps(2, params, R, T, u_old, 1)
print("step size max", np.abs(u_new - u_old).max())
def tiled(HW, HS, HN, BX=32, BY=32):
    out = u_old.copy()
    for j0 in range(1, ny - 1, BY):
        for i0 in range(1, nx - 1, BX):
            i1, j1 = min(i0 + BX, nx - 1), min(j0 + BY, ny - 1)
            ci0, ci1 = max(i0 - HW - 1, 0), min(i1 + 1, nx)
            cj0, cj1 = max(j0 - HS - 1, 0), min(j1 + HN + 1, ny)
            sub = o.solver_steps(2, params, R[cj0:cj1, ci0:ci1], T[cj0:cj1, ci0:ci1], u_old[cj0:cj1, ci0:ci1], 1)
            out[j0:j1, i0:i1] = sub[j0 - cj0:j1 - cj0, i0 - ci0:i1 - ci0]
    return out
for HW, HS, HN in [(4, 2, 1), (8, 4, 2), (12, 6, 3), (16, 8, 4), (20, 10, 6), (28, 14, 8)]:
    t = tiled(HW, HS, HN)
    d = np.abs(t.astype(np.float64) - u_new.astype(np.float64))
    print(f"HW={HW:2d} HS={HS:2d} HN={HN}: max diff {d.max():.3e}  frac differing {np.mean(d > 0):.2e}")
