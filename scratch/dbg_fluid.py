import sys, numpy as np
sys.path.insert(0, '.')
import bench
import opticalflow2d_b200 as of
size = 2048
ref = np.load(f"scratch/_big/ref_fluid_{size}.npz")
R, T = bench.make_inputs("fluid", size)
of.set_strict(False, 32)
s = of.Session((size, size), [100], 0, 5, bench.PARAMS["fluid"], nrefine=1, verbose=0, bits=32)
s.set_images(R, T); s.estimate(); tr = s.trace()["levels"][0]; s.close()
k = min(len(tr["fluid_dt"]), len(ref["fluid_dt"]))
rel_dt = np.abs(tr["fluid_dt"][:k] - ref["fluid_dt"][:k]) / np.abs(ref["fluid_dt"][:k])
rel_err = np.abs(tr["err"][:k] - ref["err"][:k]) / np.abs(ref["err"][:k] + 1e-30)
for it in range(k):
    print(it, "dt_rel %.2e err_rel %.2e" % (rel_dt[it], rel_err[it]), "dt", tr["fluid_dt"][it], ref["fluid_dt"][it])
print("minjac ours", tr["regrid_minjac"][:50])
print("minjac ref ", ref["regrid_minjac"][:50])
