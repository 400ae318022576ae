"""Generates tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/libof2d_ref{32,64}.so, i.e.
the unmodified sources of /root/reference built by oracle/Makefile).  Run in the authoring
container only:  python tests/golden/make_golden.py

Each fixture stores the synthetic inputs, the reference's outputs through its own mexFunction
(motion, warped image) and the control-flow trace captured from its mexPrintf calls.  The -m "not
gpu" tests require the C oracle to reproduce them bit for bit; the -m gpu tests hold the CUDA path to
the north-star tolerances against them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from opticalflow2d_b200 import synthetic as S  # noqa: E402
from oracle import refapi  # noqa: E402

CASES = {
    # name: (dimx, dimy, pair kwargs, reg, regparams, niter, nscales, nrefine)
    "c1_diffusion_blob": (64, 64, dict(kind="blob"), 0, [0.5], [60], 0, 1),
    "diffusion_multiscale": (80, 56, dict(kind="blob", smooth=True), 0, [0.5], [12, 16, 20], 2, 2),
    "curvature": (64, 32, dict(kind="lattice", sigma_b=6.0, smooth=True), 1, [0.25, 1.0], [25], 0, 1),
    "elastic": (72, 48, dict(kind="lattice", sigma_b=6.0, smooth=True), 2, [1.0, 0.25], [40], 0, 1),
    "thirion_compose": (64, 48, dict(kind="lattice", sigma_b=6.0, smooth=True), 3, [1, 0.25, 1.5, 1.5, 5, 0], [40], 0, 1),
    "thirion_add": (64, 48, dict(kind="lattice", sigma_b=6.0, smooth=True), 3, [1, 0.25, 1.5, 1.5, 5, 1], [25], 0, 1),
    "diffeomorphic": (64, 48, dict(kind="lattice", sigma_b=6.0, smooth=True), 4, [1, 2.0, 1.5, 1.5, 5], [25], 0, 1),
    "fluid_regrid": (96, 80, dict(kind="lattice", sigma_b=6.0, smooth=True), 5, [0.1, 0.0], [80], 0, 1),
    "fluid_multiscale": (100, 72, dict(kind="blob", smooth=True), 5, [0.25, 0.0], [25, 25], 1, 1),
}


def main():
    for name, (dimx, dimy, kw, reg, params, niter, nscales, nrefine) in CASES.items():
        R, T = S.make_pair(dimx, dimy, **kw)
        out = {"Iref": R, "Imov": T, "reg": reg, "regparams": np.asarray(params, dtype=np.float64), "niter": np.asarray(niter),
               "nscales": nscales, "nrefine": nrefine}
        for bits in (32, 64):
            ref = refapi.get("ref", bits)
            r = ref.register(R, T, reg, params, niter, nscales=nscales, nrefine=nrefine, verbose=1)
            for k, v in r.items():
                out[f"{k}_{bits}"] = v
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.startswith(("motion", "err_3", "regrid_iter"))})


if __name__ == "__main__":
    main()
