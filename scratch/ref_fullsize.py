"""Reference (compiled, CPU) at the bench's full size: iteration counts + traces, saved for comparison with the GPU run."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import bench
from oracle import refapi
m = sys.argv[1]; size = int(sys.argv[2])
lib = refapi.get("ref", 32)
R, T = bench.make_inputs(m, size)
t0 = time.time()
out = lib.register(R, T, bench.REG[m], bench.PARAMS[m], [bench.NITER[m]], nscales=0, nrefine=1, verbose=1)
print(m, size, "iterations", len(out["err"]), "regrids", len(out["regrid_iter"]), "time", time.time() - t0, flush=True)
np.savez_compressed(f"scratch/_big/ref_{m}_{size}.npz", motion=out["motion"].astype(np.float32), err=out["err"], regrid_iter=out["regrid_iter"],
                    fluid_maxabs=out["fluid_maxabs"], fluid_dt=out["fluid_dt"], regrid_minjac=out["regrid_minjac"])
