"""CPU (authoring container): the compiled reference's OWN fp32 vs fp64 builds on one bench input, max |du| over the whole field and
with a border ring excluded.  Diffeomorphic 1024^2, 50 iterations (scratch/ref_border_spread_diffeo1024.log): 1.15 px on the
outermost ring (the out-of-bounds test of Motion::accumulate, src/Motion.cpp:141-144, is a discontinuity: a pixel whose
composed position is within rounding of the image edge either keeps its old value or takes the interpolated one), 3e-2 one pixel
in, 6e-4 eight pixels in.  usage: python scratch/ref_border_spread.py <size> <method> <niter>"""
import sys, numpy as np, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import refapi
size = int(sys.argv[1]); m = sys.argv[2]; niter = int(sys.argv[3])
R, T = bench.make_inputs(m, size)
out = {}
for bits in (32, 64):
    lib = refapi.get("ref", bits)
    t0 = time.time()
    o = lib.register(R, T, bench.REG[m], bench.PARAMS[m], [niter], nscales=0, nrefine=1, verbose=1)
    out[bits] = o["motion"].astype(np.float64)
    print(bits, len(o["err"]), time.time() - t0, flush=True)
d = np.abs(out[32] - out[64]).max(axis=2)
print("max", d.max(), "argmax", np.unravel_index(d.argmax(), d.shape))
for k in (0, 1, 2, 4, 8):
    inner = d[k:size - k, k:size - k] if k else d
    print("excluding border ring", k, "max", inner.max(), "n>1e-3", int((inner > 1e-3).sum()))
print("rows with >1e-3:", sorted(set(np.argwhere(d > 1e-3)[:, 0].tolist()))[:20], "cols:", sorted(set(np.argwhere(d > 1e-3)[:, 1].tolist()))[:20])
