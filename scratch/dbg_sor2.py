import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'scratch')
import numpy as np, torch
from gpu_common import *
dev = device()
bits = 32; f32 = np.float32
dimx, dimy = 64, 48
orc = oracle(bits)
R, T = pair(dimx, dimy, "lattice", smooth=True, sigma_b=6.0)
R, T = R.astype(NP[bits]), T.astype(NP[bits])
g, it = orc.derivatives(R, T)
u0 = S.random_motion(dimx, dimy, 0.3, 21, True).astype(NP[bits])
def run(u, g, it, label):
    # oracle via solver_steps needs images; emulate through a direct python sweep instead
    from dbg_sor_emu import sweep
    want = sweep(u, g, it, f32(1.0), f32(0.25), f32(0.66))
    d_u = to_dev(u)
    dev.call("elastic_step", TD[bits], dimx, dimy, 1, d_u, to_dev(g), to_dev(it), f32(1.0), f32(0.25), f32(0.66))
    got = d_u.cpu().numpy()
    d = np.abs(got-want)
    print(label, "band0 max", d[:, 1:33].max(), "band1 max", d[:, 33:63].max(), "first bad in band1 (j,i):", next(((j,i) for i in range(33,63) for j in range(1,47) if d[j,i].max()>0), None))
    return got, want
run(np.zeros_like(u0), g, it, "A u0=0       ")
run(u0, np.zeros_like(g), np.zeros_like(it), "B force=0    ")
u1 = u0.copy(); u1[:, 33:] = 0
run(u1, np.zeros_like(g), np.zeros_like(it), "C force=0, band1 cols zero ")
u2 = u0.copy(); u2[:, :33] = 0
got, want = run(u2, np.zeros_like(g), np.zeros_like(it), "D force=0, band0 cols zero ")
print(got[1:4,33], want[1:4,33])
