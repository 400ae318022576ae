"""Elastic at 2048^2: time and result of the tiled sweep for several halo truncations (OF2D_SOR_EPS_LOG2)."""
import os, time, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import opticalflow2d_b200 as of
import bench
R, T = bench.make_inputs("elastic", 2048)
ref = None
for eps in (-40, -34, -30, -26):
    os.environ["OF2D_SOR_EPS_LOG2"] = str(eps)
    with of.Session((2048, 2048), [50], 0, of.ELASTIC, [1.0, 0.25], nrefine=1, verbose=0, bits=32) as s:
        s.set_images(R, T)
        s.estimate()
        ts = []
        for _ in range(3):
            s.reset(); t0 = time.perf_counter(); s.estimate(); ts.append(time.perf_counter() - t0)
        m = s.motion()
    if ref is None: ref = m
    print(f"eps 2^{eps}: {1e3*min(ts):.3f} ms / 50 iterations, max|du| vs 2^-40 = {float(np.max(np.abs(m-ref))):.3e} px", flush=True)
