import numpy as np
f32=np.float32
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
