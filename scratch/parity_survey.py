"""GPU: every fixture of tests/golden (configs at their own sizes) in fast and strict mode -> max|du|, traces, SSD."""
import glob, os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import opticalflow2d_b200 as of
import make_golden_configs as G

pat = sys.argv[1:] or ["full2048_fluid_c", "full2048f64_", "c2_", "c3_", "c5_", "demo_"]
modes = os.environ.get("MODES", "relaxed,exact,strict").split(",")
cases = G.cases()
rows = []
for name, spec in cases.items():
    if not any(name.startswith(p) for p in pat):
        continue
    path = os.path.join(G.GOLD, name + ".npz")
    if not os.path.exists(path):
        continue
    g = np.load(path)
    R, T = G.make_inputs(spec)
    dimy, dimx = R.shape
    bits = int(g["bits"])
    for mode in modes:
        if mode == "strict" and spec["method"] in ("elastic", "fluid") and dimx >= 2048 and os.environ.get("STRICT_BIG", "1") == "0":
            continue
        of.set_math(mode, bits)
        t0 = time.time()
        with of.Session((dimx, dimy), [int(v) for v in g["niter"]], int(g["nscales"]), int(g["reg"]), list(g["regparams"]), nrefine=1, verbose=0, bits=bits) as s:
            s.set_images(R, T)
            s.estimate()
            mo = s.motion()
            tr = s.trace()
            warped = s.warp(T)
        st, off = int(g["stride"]), int(g["offset"])
        du = float(np.abs(mo[off::st, off::st] - g["sample"].astype(np.float64)).max())
        err = np.concatenate([l["err"] for l in tr["levels"]]) if tr["levels"] else np.zeros(0)
        rg = np.concatenate([l["regrid_iter"] for l in tr["levels"]]) if tr["levels"] else np.zeros(0)
        n = min(len(err), len(g["err"]))
        rel = float((np.abs(err[:n] - g["err"][:n]) / np.maximum(np.abs(g["err"][:n]), 1e-12)).max()) if n else 0.0
        ssd1 = float(((warped - R) ** 2).sum())
        row = dict(case=name, mode=mode, bits=bits, du=du, iters=int(tr["total_iterations"]), ref_iters=int(len(g["err"])),
                   regrids=int(len(rg)), ref_regrids=int(len(g["regrid_iter"])), regrid_same=bool(np.array_equal(rg.astype(int), g["regrid_iter"].astype(int))),
                   err_rel=rel, ssd_rel=abs(ssd1 - float(g["ssd1"])) / float(g["ssd1"]), mean_d=float(np.abs(mo.mean(axis=(0, 1)) - g["mean"]).max()), sec=time.time() - t0)
        rows.append(row)
        print(json.dumps(row), flush=True)
of.set_math("relaxed", 32); of.set_math("relaxed", 64)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", os.environ.get("SURVEY_OUT", "parity_survey.json")), "w"), indent=1)
