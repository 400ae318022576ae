import sys, numpy as np
sys.path.insert(0, '.')
import opticalflow2d_b200 as of
from opticalflow2d_b200 import synthetic as S
dimx, dimy = 200, 300
R, T = S.make_pair(dimx, dimy, "lattice", shift=(1.5, -0.75), sigma_b=6.0)
for bits in (64, 32):
    for niter in (1, 2, 5, 20):
        mo = {}
        for strict in (False, True):
            of.set_strict(strict, bits)
            with of.Session((dimx, dimy), [niter], 0, of.ELASTIC, [1.0, 0.25], bits=bits) as s:
                s.set_images(R, T); s.estimate(); mo[strict] = s.motion()
        d = np.abs(mo[True] - mo[False]).max(axis=2)
        j, i = np.unravel_index(np.argmax(d), d.shape)
        rows = np.nonzero(d.max(axis=1) > 0.1 * d.max())[0]; cols = np.nonzero(d.max(axis=0) > 0.1 * d.max())[0]
        print(bits, niter, "max", d.max(), "at j,i", j, i, "rows", rows[:12], "cols", cols[:12], "scale", np.abs(mo[True]).max(), flush=True)
