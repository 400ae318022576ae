"""GPU box with >= 2 GPUs: one process drives all of them through of2d_batch_create_multi (one host thread + context + copy streams
per device); result == the single-device batch bit for bit (exact engine), and pairs/s of the streamed protocol for 1 vs N devices."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import opticalflow2d_b200 as of
from opticalflow2d_b200 import synthetic as S
import bench
ndev = torch.cuda.device_count()
px, n = 512, 256 * ndev
R = np.empty((n, px, px)); T = np.empty((n, px, px))
for k in range(n):
    R[k], T[k] = S.batch_pair(k, px, px)
pr = torch.from_numpy(R).pin_memory(); pt = torch.from_numpy(T).pin_memory(); po = torch.empty((n, 2, px, px), dtype=torch.float64).pin_memory()
out = {"devices": ndev, "pairs": n}
for m in ("thirion", "fluid"):
    res = {}
    for devs in ([0], list(range(ndev))):
        with of.Batch((px, px), n, bench.BATCH_NITER[m], bench.REG[m], bench.PARAMS[m], wave=64, devices=devs) as b:
            b.register_raw(pr.data_ptr(), pt.data_ptr(), po.data_ptr())     # warm-up
            t0 = time.perf_counter()
            b.register_raw(pr.data_ptr(), pt.data_ptr(), po.data_ptr())
            dt = time.perf_counter() - t0
            res[len(devs)] = {"pairs_per_s": n / dt, "shards": b.shards(), "motion": po.numpy().copy(), "its": b.iterations()[0].copy()}
    same = bool(np.array_equal(res[1]["motion"], res[ndev]["motion"]) and np.array_equal(res[1]["its"], res[ndev]["its"]))
    out[m] = {"pairs_per_s_1dev": res[1]["pairs_per_s"], f"pairs_per_s_{ndev}dev": res[ndev]["pairs_per_s"], "speedup": res[ndev]["pairs_per_s"] / res[1]["pairs_per_s"],
              "identical_to_single_device": same, "shards": res[ndev]["shards"]}
print(json.dumps(out), flush=True)
