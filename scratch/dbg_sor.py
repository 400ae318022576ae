import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from gpu_common import *
dev = device()
bits = 32
f32 = np.float32
def sweep(u, g, it, mu, la, om, stale_west_col=None, u_old=None):
    u = u.copy(); ny, nx, _ = u.shape
    b = np.empty_like(u)
    s = it + u[...,0]*g[...,0] + u[...,1]*g[...,1]
    b[...,0] = g[...,0]*s; b[...,1] = g[...,1]*s
    ck = f32(1.0)-om; cr = om/(f32(-6)*mu - f32(2)*la); mupl = mu+la
    def X(i,j,c, cur_i):
        if stale_west_col is not None and i == stale_west_col and cur_i == stale_west_col+1: return u_old[j,i,c]
        return u[j,i,c]
    for i in range(1,nx-1):
        for j in range(1,ny-1):
            nx_ = ck*u[j,i,0] + cr*(b[j,i,0] - mu*(X(i+1,j,0,i)+X(i-1,j,0,i)+X(i,j+1,0,i)+X(i,j-1,0,i)) - mupl*(X(i+1,j,0,i)+X(i-1,j,0,i)+f32(0.25)*(X(i+1,j+1,1,i)-X(i-1,j+1,1,i)-X(i+1,j-1,1,i)+X(i-1,j-1,1,i))))
            ny_ = ck*u[j,i,1] + cr*(b[j,i,1] - mu*(X(i+1,j,1,i)+X(i-1,j,1,i)+X(i,j+1,1,i)+X(i,j-1,1,i)) - mupl*(X(i+1,j,1,i)+X(i-1,j,1,i)+f32(0.25)*(X(i+1,j+1,0,i)-X(i-1,j+1,0,i)-X(i+1,j-1,0,i)+X(i-1,j-1,0,i))))
            u[j,i,0] = nx_; u[j,i,1] = ny_
    return u
dimx, dimy = 64, 48
orc = oracle(bits)
R, T = pair(dimx, dimy, "lattice", smooth=True, sigma_b=6.0)
R, T = R.astype(NP[bits]), T.astype(NP[bits])
g, it = orc.derivatives(R, T)
u0 = S.random_motion(dimx, dimy, 0.3, 21, True).astype(NP[bits])
mu, la, om = f32(1.0), f32(0.25), f32(0.66)
want = orc.solver_steps(2, [1.0,0.25,0.66], R, T, u0, 1)
emu = sweep(u0, g, it, mu, la, om)
print("numpy emu vs oracle", np.abs(emu-want).max())
emu_stale = sweep(u0, g, it, mu, la, om, stale_west_col=32, u_old=u0)
d_u = to_dev(u0)
dev.call("elastic_step", TD[bits], dimx, dimy, 1, d_u, to_dev(g), to_dev(it), mu, la, om)
got = d_u.cpu().numpy()
print("gpu vs oracle", np.abs(got-want).max(), " gpu vs stale-west emu", np.abs(got-emu_stale).max())
print("col 32 gpu==oracle:", np.array_equal(got[:,32], want[:,32]), "col 33 row1:", got[1,33], want[1,33], emu_stale[1,33])
# repeat the same call on fresh data (second launch)
d_u = to_dev(u0)
dev.call("elastic_step", TD[bits], dimx, dimy, 1, d_u, to_dev(g), to_dev(it), mu, la, om)
print("second launch gpu vs oracle", np.abs(d_u.cpu().numpy()-want).max())
