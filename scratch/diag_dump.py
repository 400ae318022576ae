import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import opticalflow2d_b200 as of
tag = sys.argv[1]
m, size = "diffeomorphic", 2048
R, T = bench.make_inputs(m, size)
of.set_math("relaxed", 32)
for niter in (11, 12, 14, 17, 20, 25, 30, 40):
    with of.Session((size, size), [niter], 0, bench.REG[m], bench.PARAMS[m], nrefine=1, verbose=0, bits=32) as s:
        s.set_images(R, T); s.estimate()
        mo = s.motion(); tr = s.trace()["levels"][0]
    np.save(f"/tmp/diag_{tag}_{niter}.npy", mo.astype(np.float32))
    print(tag, niter, tr.get("nsq"), flush=True)
