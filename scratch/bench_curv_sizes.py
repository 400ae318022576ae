"""GPU: Curvature iteration rate per pixel at power-of-two and non-power-of-two sizes (Bluestein), fp32 fields."""
import os, sys, json, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import opticalflow2d_b200 as of
from opticalflow2d_b200 import synthetic as S
of.set_stream(torch.cuda.current_stream().cuda_stream, 32)
for dimx, dimy in [(256, 256), (278, 256), (512, 512), (600, 360), (1024, 1024), (1000, 1000), (2048, 2048), (1800, 1500)]:
    R, T = S.make_pair(dimx, dimy, "lattice", shift=(1.5, -0.75))
    with of.Session((dimx, dimy), [30], 0, of.CURVATURE, [0.25, 1.0], nrefine=1, verbose=0, bits=32) as s:
        s.set_images(R, T)
        for rep in range(3):
            s.reset()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); s.estimate(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        it = s.trace()["total_iterations"]
    print(json.dumps({"size": [dimx, dimy], "ms": ms, "iterations": it, "mpix_iter_s": dimx * dimy * it / ms / 1e3}), flush=True)
