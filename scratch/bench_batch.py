"""Batch leg only: pairs/s (resident) of 512 pairs of 512^2 for several wave sizes.  python scratch/bench_batch.py thirion 64,128,148,222,256"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import opticalflow2d_b200 as of
m = sys.argv[1] if len(sys.argv) > 1 else "thirion"
waves = [int(w) for w in (sys.argv[2] if len(sys.argv) > 2 else "64,128,148,222,256").split(",")]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 512
bp = bench.BATCH_PX
Rb, Tb = bench.make_batch_inputs(0, B, bp)
prb = torch.from_numpy(Rb).pin_memory(); ptb = torch.from_numpy(Tb).pin_memory()
pob = torch.empty((B, 2, bp, bp), dtype=torch.float64).pin_memory()
for w in waves:
    bt = of.Batch((bp, bp), B, bench.BATCH_NITER[m], bench.REG[m], bench.PARAMS[m], nrefine=1, wave=min(B, w), bits=32)
    bt.set_images_raw(prb.data_ptr(), ptb.data_ptr())
    out = {}
    for leg in ("resident", "e2e"):
        f = bt.estimate if leg == "resident" else (lambda: bt.register_raw(prb.data_ptr(), ptb.data_ptr(), pob.data_ptr()))
        f(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); f(); e1.record(); torch.cuda.synchronize()
        out[leg] = round(2 * B / (e0.elapsed_time(e1) * 1e-3), 1)
    its, _ = bt.iterations()
    print(m, "wave", w, out, "mean it", float(its.mean()), flush=True)
    bt.close()
