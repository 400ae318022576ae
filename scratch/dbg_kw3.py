import sys, numpy as np
sys.path.insert(0, ".")
import opticalflow2d_b200 as of
from opticalflow2d_b200 import synthetic as S
dimx, dimy = int(sys.argv[1]) if len(sys.argv) > 1 else 160, int(sys.argv[2]) if len(sys.argv) > 2 else 96
kw = int(sys.argv[3]) if len(sys.argv) > 3 else 3
R, T = S.make_pair(dimx, dimy, "lattice", shift=(1.5, -0.75), smooth=True, sigma_b=8.0)
of.set_math("relaxed", 32)
with of.Session((dimx, dimy), [10], 0, of.THIRION, [1.0, 0.5, 1.0, 0.8, kw, 0], nrefine=1, verbose=0, bits=32) as s:
    s.set_images(R, T)
    s.estimate()
    m = s.motion()
print("ok", float(np.abs(m).max()))
