"""Golden vectors at BASELINE.json's full size (2048 x 2048), from the compiled reference (oracle/_ref).

Runs in the authoring container only (needs oracle/_ref, i.e. /root/reference; minutes of CPU per method).  The
full motion field (32 MiB) does not belong in git: the fixture keeps the reference's motion SAMPLED on a 64 x 64
lattice of pixels (every 32nd row / column, offset 5 so that blob centres, flanks and background are all hit),
its iteration count, Logger error series, regrid iterations and per-plane mean / mean-square -- enough to pin
the GPU path at full size (tests/test_fullsize_gpu.py).

    python tests/golden/make_golden_full.py [method ...]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload definitions: inputs, parameters)
from oracle import refapi  # noqa: E402

SIZE = 2048
NITER = {"diffusion": 50, "curvature": 50, "elastic": 50, "thirion": 50, "diffeomorphic": 50, "fluid": 40}
STRIDE, OFFSET = 32, 5


def main(methods):
    lib = refapi.get("ref", 32)
    for m in methods:
        R, T = bench.make_inputs(m, SIZE)
        out = lib.register(R, T, bench.REG[m], bench.PARAMS[m], [NITER[m]], nscales=0, nrefine=1, verbose=1)
        mo = out["motion"]
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"full{SIZE}_{m}.npz"),
                            method=m, size=SIZE, reg=bench.REG[m], regparams=np.asarray(bench.PARAMS[m], dtype=np.float64), niter=NITER[m],
                            stride=STRIDE, offset=OFFSET, sample=mo[OFFSET::STRIDE, OFFSET::STRIDE].astype(np.float32),
                            mean=mo.mean(axis=(0, 1)), meansq=(mo.astype(np.float64) ** 2).mean(axis=(0, 1)),
                            err=out["err"], regrid_iter=out["regrid_iter"], fluid_dt=out["fluid_dt"])
        print(m, "iterations", len(out["err"]), "regrids", len(out["regrid_iter"]), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or list(NITER))
