"""Golden vectors for BASELINE.json's configurations at their own sizes, from the compiled reference (oracle/_ref).

Runs in the authoring container only (needs oracle/_ref, i.e. /root/reference); every case is one process, the
cases run in parallel on the host cores (minutes in total).  As in make_golden_full.py a fixture keeps the
reference's motion SAMPLED on a lattice of pixels, the iteration count, the Logger error series, the regrid
iterations / Fluid time steps and the per-plane mean / mean-square of the motion.

    python tests/golden/make_golden_configs.py [case ...]      (no argument: all cases that do not exist yet)

Cases
  full2048_fluid_c{40,60,80,100}      config 4, Fluid at several iteration caps (fp32): where does fast mode leave the
                                       reference's regrid trace?  (the bench cap is the largest one that is pinned)
  full2048f64_<method>[_c<cap>]       config 4 in fp64 (libof2d_ref64.so = the reference after s/float/double/)
  c2_thirion_512                      config 2: Thirion, 512^2, smooth deformation, sigma 1.5, 200 iterations
  c3_diffeomorphic_1024               config 3: Diffeomorphic, 1024^2, sigma_x = 2 (squarings active), 100 iterations
  c5[f64]_<method>[_c60]_pair<k>      config 5: batch_pair(k) of 512^2, Thirion and Fluid, cap 100 (and Fluid at cap 60, fp32 / fp64), k = 0..7
  demo_fluid_278x256                  the demo's call (test_opticalflow2d.m:23-38): Fluid, nscales 1, niter [25 25],
                                       alpha [0.25 0], on a 256^2 slice padded by 11 rows on both sides (278 x 256)
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload definitions: inputs, parameters)
from opticalflow2d_b200 import synthetic as S  # noqa: E402
from oracle import refapi  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
FLUID_CAPS = (40, 60, 80, 100)
F64_NITER = {"diffusion": 50, "curvature": 50, "elastic": 50, "thirion": 50, "diffeomorphic": 50}
C5_PAIRS = range(8)


def demo_pair():
    """A 256 x 256 lattice slice normalised to [0, 1] and padded by 11 replicated rows on both sides of the first
    (fast) MATLAB dimension, as test_opticalflow2d.m:14-18 does with the DIR-Lab slice: 278 x 256."""
    R, T = S.make_pair(256, 256, kind="lattice", shift=(1.5, -0.75), smooth=True, sigma_b=6.0)
    out = []
    for a in (R, T):
        a = (a - a.min()) / (a.max() - a.min())
        out.append(np.ascontiguousarray(np.pad(a, ((0, 0), (11, 11)), mode="edge")))   # x (fast index) is MATLAB's first dimension
    return out[0], out[1]


def cases():
    c = {}
    for cap in FLUID_CAPS:
        c[f"full2048_fluid_c{cap}"] = dict(method="fluid", bits=32, niter=[cap], inputs=("bench", 2048))
        c[f"full2048f64_fluid_c{cap}"] = dict(method="fluid", bits=64, niter=[cap], inputs=("bench", 2048))
    for m, n in F64_NITER.items():
        c[f"full2048f64_{m}"] = dict(method=m, bits=64, niter=[n], inputs=("bench", 2048))
    c["c2_thirion_512"] = dict(method="thirion", bits=32, niter=[200], inputs=("smooth", 512))
    c["c3_diffeomorphic_1024"] = dict(method="diffeomorphic", bits=32, niter=[100], inputs=("lattice", 1024))
    for k in C5_PAIRS:
        for m in ("thirion", "fluid"):
            c[f"c5_{m}_pair{k}"] = dict(method=m, bits=32, niter=[100], inputs=("batch", k))
        # Fluid amplifies rounding noise (regridding, explicit Euler): the fp64 build of the reference on the same pairs measures the
        # reference's own sensitivity, and the cap-60 runs are where it still agrees with its fp64 build to 1e-3 px
        c[f"c5f64_fluid_pair{k}"] = dict(method="fluid", bits=64, niter=[100], inputs=("batch", k))
        c[f"c5_fluid_c60_pair{k}"] = dict(method="fluid", bits=32, niter=[60], inputs=("batch", k))
        c[f"c5f64_fluid_c60_pair{k}"] = dict(method="fluid", bits=64, niter=[60], inputs=("batch", k))
    c["demo_fluid_278x256"] = dict(method="fluid", bits=32, niter=[25, 25], nscales=1, params=[0.25, 0.0], inputs=("demo", 0))
    return c


def make_inputs(spec):
    kind, arg = spec["inputs"]
    if kind == "bench":
        return bench.make_inputs(spec["method"], arg)
    if kind == "smooth":
        return S.make_pair(arg, arg, kind="lattice", shift=(1.5, -0.75), smooth=True, sigma_b=8.0)
    if kind == "lattice":
        return S.make_pair(arg, arg, kind="lattice", shift=(1.5, -0.75), smooth=False, sigma_b=8.0)
    if kind == "batch":
        return S.batch_pair(arg, 512, 512)
    if kind == "demo":
        return demo_pair()
    raise ValueError(kind)


def run_case(name):
    spec = cases()[name]
    lib = refapi.get("ref", spec["bits"])
    R, T = make_inputs(spec)
    m = spec["method"]
    params = spec.get("params", bench.PARAMS[m])
    nscales = spec.get("nscales", 0)
    t0 = time.time()
    out = lib.register(R, T, bench.REG[m], params, spec["niter"], nscales=nscales, nrefine=1, verbose=1)
    mo = out["motion"]
    dimy, dimx = R.shape
    stride = 32 if max(dimx, dimy) >= 2048 else 16 if max(dimx, dimy) >= 1024 else 8
    offset = 5
    keep = np.float64 if spec["bits"] == 64 else np.float32
    np.savez_compressed(os.path.join(GOLD, name + ".npz"),
                        method=m, bits=spec["bits"], dims=np.asarray([dimx, dimy]), reg=bench.REG[m], regparams=np.asarray(params, dtype=np.float64),
                        niter=np.asarray(spec["niter"]), nscales=nscales, inputs=np.asarray([str(spec["inputs"][0]), str(spec["inputs"][1])]),
                        stride=stride, offset=offset, sample=mo[offset::stride, offset::stride].astype(keep),
                        mean=mo.mean(axis=(0, 1)), meansq=(mo.astype(np.float64) ** 2).mean(axis=(0, 1)),
                        maxabs=np.abs(mo).max(axis=(0, 1)),
                        ssd0=float(((T - R) ** 2).sum()), ssd1=float(((out["warped"] - R) ** 2).sum()),
                        err=out["err"], err_iter=out["err_iter"], regrid_iter=out["regrid_iter"], regrid_minjac=out["regrid_minjac"],
                        fluid_dt=out["fluid_dt"], fluid_maxabs=out["fluid_maxabs"])
    return name, len(out["err"]), len(out["regrid_iter"]), time.time() - t0


def main(argv):
    names = argv or [n for n in cases() if not os.path.exists(os.path.join(GOLD, n + ".npz"))]
    # longest first
    order = sorted(names, key=lambda n: -(cases()[n]["niter"][0] * (cases()[n]["inputs"][1] if cases()[n]["inputs"][0] == "bench" else 300) ** 2))
    with mp.get_context("fork").Pool(min(len(order), os.cpu_count() or 1) or 1) as pool:
        for name, it, rg, sec in pool.imap_unordered(run_case, order):
            print(f"{name}: iterations {it} regrids {rg} ({sec:.0f} s)", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
