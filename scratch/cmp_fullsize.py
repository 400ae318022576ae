"""GPU vs compiled reference at the bench's full size (reference results precomputed on CPU into scratch/_big)."""
import sys, json, numpy as np
sys.path.insert(0, '.')
import bench
import opticalflow2d_b200 as of
size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
methods = sys.argv[2].split(",") if len(sys.argv) > 2 else bench.METHODS
res = {}
for m in methods:
    ref = np.load(f"scratch/_big/ref_{m}_{size}.npz")
    R, T = bench.make_inputs(m, size)
    for strict in (False, True):
        if strict and m in ("elastic", "fluid") and size > 512:
            continue   # the exact wavefront is slow at this size
        of.set_strict(strict, 32)
        s = of.Session((size, size), [bench.NITER[m]], 0, bench.REG[m], bench.PARAMS[m], nrefine=1, verbose=0, bits=32)
        s.set_images(R, T); s.estimate()
        mo = s.motion(); tr = s.trace()["levels"][0]; s.close()
        err = tr["err"]; n = min(len(err), len(ref["err"]))
        rel = np.abs(err[:n] - ref["err"][:n]) / np.maximum(np.abs(ref["err"][:n]), 1e-30)
        bad = np.nonzero(rel > 5e-4)[0]
        out = {"iterations": int(tr["iterations"]), "ref_iterations": int(len(ref["err"])), "max_du": float(np.abs(mo - ref["motion"]).max()),
               "first_err_mismatch": int(bad[0]) if len(bad) else -1, "max_err_rel": float(rel.max()),
               "regrid": [int(x) for x in tr["regrid_iter"]][:60], "ref_regrid": [int(x) for x in ref["regrid_iter"]][:60]}
        if m == "fluid":
            k = min(len(tr["fluid_dt"]), len(ref["fluid_dt"]))
            out["dt_rel_max"] = float(np.max(np.abs(tr["fluid_dt"][:k] - ref["fluid_dt"][:k]) / np.abs(ref["fluid_dt"][:k])))
            out["err_tail"] = [float(x) for x in err[-6:]]; out["ref_err_tail"] = [float(x) for x in ref["err"][-6:]]
        res[f"{m}/{'strict' if strict else 'fast'}"] = out
        print(m, "strict" if strict else "fast", json.dumps(out), flush=True)
json.dump(res, open(f"gpurun_out/cmp_fullsize_{size}.json", "w"), indent=1)
