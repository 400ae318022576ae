"""Largest relative difference of the Logger error series against the compiled reference's, per full-size fixture (2048^2, fp32)."""
import glob, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import opticalflow2d_b200 as of
for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "full2048_*.npz"))):
    name = os.path.basename(path)[9:-4]
    if name not in bench.REG: continue
    g = np.load(path)
    method, size, niter = str(g["method"]), int(g["size"]), int(g["niter"])
    R, T = bench.make_inputs(method, size)
    with of.Session((size, size), [niter], 0, bench.REG[method], bench.PARAMS[method], nrefine=1, verbose=0, bits=32) as s:
        s.set_images(R, T); s.estimate(); tr = s.trace()["levels"][0]
    rel = np.abs(np.asarray(tr["err"]) - g["err"]) / np.maximum(np.abs(g["err"]), 1e-12)
    print(f"{method:14s} rel.max={rel.max():.3e} at it {int(rel.argmax())} (err there {g['err'][int(rel.argmax())]:.4e}); median {np.median(rel):.2e}; min err {g['err'].min():.3e}", flush=True)
