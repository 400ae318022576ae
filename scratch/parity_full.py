"""GPU: the six 2048^2 fp32 fixtures (tests/golden/full2048_<m>.npz) at the arithmetic level given by OF2D_MATH / OF2D_FUSED:
max |du| on the samples, iteration counts, nsquares trace differences vs the exact engine."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import opticalflow2d_b200 as of
methods = sys.argv[1:] or bench.METHODS
for m in methods:
    g = np.load(os.path.join(ROOT, "tests", "golden", f"full2048_{m}.npz"))
    R, T = bench.make_inputs(m, 2048)
    res = {}
    for level in ("exact", "relaxed"):
        of.set_math(level, 32)
        with of.Session((2048, 2048), [int(g["niter"])], 0, bench.REG[m], bench.PARAMS[m], nrefine=1, verbose=0, bits=32) as s:
            s.set_images(R, T); s.estimate()
            mo = s.motion(); tr = s.trace()["levels"][0]
        st, off = int(g["stride"]), int(g["offset"])
        res[level] = (mo, tr)
        print(json.dumps({"method": m, "level": level, "fused": os.environ.get("OF2D_FUSED", "1"), "du": float(np.abs(mo[off::st, off::st] - g["sample"]).max()),
                          "iters": tr["iterations"], "ref_iters": int(len(g["err"])), "nsq": [int(v) for v in tr.get("nsq", [])][:60] if m == "diffeomorphic" else None}), flush=True)
    d = np.abs(res["exact"][0] - res["relaxed"][0])
    k = np.unravel_index(d.argmax(), d.shape)
    print(json.dumps({"method": m, "relaxed_vs_exact_max": float(d.max()), "at": [int(v) for v in k], "p99.9": float(np.quantile(d, 0.999)), "mean": float(d.mean())}), flush=True)
