import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import opticalflow2d_b200 as of
m = sys.argv[1] if len(sys.argv) > 1 else "diffeomorphic"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
R, T = bench.make_inputs(m, size)
for niter in (1, 2, 3, 5, 10, 50):
    res = {}
    for level in ("exact", "relaxed"):
        of.set_math(level, 32)
        with of.Session((size, size), [niter], 0, bench.REG[m], bench.PARAMS[m], nrefine=1, verbose=0, bits=32) as s:
            s.set_images(R, T); s.estimate()
            res[level] = s.motion()
    d = np.abs(res["exact"] - res["relaxed"]).max(axis=2)
    bad = np.argwhere(d > 1e-4)
    print(niter, "max", float(d.max()), "n>1e-4", len(bad), "first", bad[:12].tolist(), "rows", sorted(set(bad[:, 0].tolist()))[:10], "cols", sorted(set(bad[:, 1].tolist()))[:10], flush=True)
    if len(bad):
        j, i = bad[0]
        print("   exact", res["exact"][j, i], "relaxed", res["relaxed"][j, i], flush=True)
