import sys, numpy as np
sys.path.insert(0, '.')
import bench
import opticalflow2d_b200 as of
for size in (256, 1024, 2048):
    R, T = bench.make_inputs("thirion", size)
    for niter in (1, 2, 3, 5, 10, 50):
        mo = {}
        for strict in (False, True):
            of.set_strict(strict, 32)
            s = of.Session((size, size), [niter], 0, 3, bench.PARAMS["thirion"], nrefine=1, verbose=0, bits=32)
            s.set_images(R, T); s.estimate(); mo[strict] = s.motion(); s.close()
        d = np.abs(mo[True] - mo[False])
        k = np.unravel_index(np.argmax(d), d.shape)
        print(size, niter, "max diff", d.max(), "at (j,i,c)", k, "count>1e-4", int((d > 1e-4).sum()), "vals", mo[True][k], mo[False][k], flush=True)
