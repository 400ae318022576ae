#!/usr/bin/env python
"""bench.py -- OpticalFlow2d registration-solve benchmark (B200, one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 2048]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json metric "Mpixel.iter/s ... per method at 2048^2"): one STEP runs the complete
`estimate_motion()` of every registration method (Diffusion/Horn-Schunck, Curvature, Elastic, Thirion
Demons, Diffeomorphic Demons, Fluid) on a synthetic 2048x2048 lattice pair, fp32,
nscales=0, nrefine=1, iteration caps NITER below, cold start.  value = sum(pixels x iterations
executed) / sum(device time); `methods` carries the per-method numbers.  N > 1: every rank runs the
same step on its own pair (weak scaling, no collective in the solve); value = sum over ranks / max time.

`value`  : images resident in HBM, CUDA-event time around estimate_motion() only.
`e2e`    : the same step through the C-ABI (include/of2d_host.h) with HOST double buffers, H2D of both images of every method and
           D2H of every planar double motion inside the timed region: one of2d_sessions_register call for the six methods (copies
           of the neighbouring methods under each solve); `e2e.sequential` = the three session calls per method, nothing overlapped.
`roofline`: the dominant kernel of the step, timed alone on L2-flushed 2048^2 inputs.
`cpu_baseline` / `--impl reference`: the reference's own sources (oracle/_ref, compiled unchanged) on the host cores.
           A full step at the GPU arm's caps takes the reference about a minute, so each CPU step is a bounded sample:
           every method at 3 and at 6 iterations -> set-up cost and per-iteration cost per method; `value` is the work
           of the GPU arm's step (same caps) over set-up + cap x per-iteration, i.e. what the full step would take.
`parity`  : the result of the timed configuration compared with the fixtures sampled from the compiled reference at the
           same size and caps (tests/golden/full2048_*.npz): iteration counts, regrid traces, max |du| on the samples.
`batch`   : BASELINE.json configs[4], 4096 pairs of 512^2 sharded over the ranks; the LAST key of the line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METHODS = ["diffusion", "curvature", "elastic", "thirion", "diffeomorphic", "fluid"]
REG = {"diffusion": 0, "curvature": 1, "elastic": 2, "thirion": 3, "diffeomorphic": 4, "fluid": 5}
# SURVEY.md 8(d) parameters
PARAMS = {
    "diffusion": [0.5],
    "curvature": [0.25, 1.0],
    "elastic": [1.0, 0.25],
    "thirion": [1.0, 0.25, 1.5, 1.5, 5, 0],
    "diffeomorphic": [1.0, 2.0, 1.5, 1.5, 5],
    "fluid": [0.1, 0.0],
}
# Fluid: 40 is the largest pinned cap at which the REFERENCE's own float and double builds still agree to the north-star tolerances
# on this input (1.6e-4 px / 8e-5 SSD at 40; 4.7e-4 px / 1.8e-3 SSD at 60; 4.6e-2 px at 80; 0.2 px and different regrid traces at 100:
# tests/golden/full2048[f64]_fluid_c*.npz) -- beyond it the reference's result is set by rounding noise, so there is nothing to be equal to
NITER = {"diffusion": 50, "curvature": 50, "elastic": 50, "thirion": 50, "diffeomorphic": 50, "fluid": 40}
SIGMA_B = {"fluid": 6.0}
# algorithmic bytes per pixel per iteration, fp32 (SURVEY.md 8d / DESIGN.md)
BYTES_PER_PX_ITER = {"diffusion": 28, "curvature": 92, "elastic": 28, "thirion": 48, "diffeomorphic": 48, "fluid": 60}
# SURVEY.md 8(d)'s own table where it differs from the layout built here: Curvature with a 4-byte spectrum (here: double, as the reference)
BYTES_PER_PX_ITER_SURVEY = {"diffusion": 28, "curvature": 60, "elastic": 28, "thirion": 48, "diffeomorphic": 48, "fluid": 60}


def make_inputs(method: str, size: int):
    from opticalflow2d_b200 import synthetic as S
    return S.make_pair(size, size, kind="lattice", shift=(1.5, -0.75), smooth=False, sigma_b=SIGMA_B.get(method, 8.0))


# --------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the compiled reference on host cores, bounded sample
# --------------------------------------------------------------------------------------------------
def _cpu_one(args):
    """One method of the step on the compiled reference: runs at n1 = 3 (its minimum) and n2 = 6 iterations; from the two
    wall times the set-up (warp + derivatives + final compose) and the per-iteration cost follow."""
    method, size, replica = args
    from oracle import refapi
    kind = "ref" if refapi.available("ref", 32) else "oracle"
    lib = refapi.get(kind, 32)
    R, T = make_inputs(method, size)
    ts, its = [], []
    for niter in CPU_SAMPLE_ITERS:
        t0 = time.perf_counter()
        out = lib.register(R, T, REG[method], PARAMS[method], [niter], nscales=0, nrefine=1, verbose=1)
        ts.append(time.perf_counter() - t0)
        its.append(len(out["err"]))
    return method, its, ts, kind


CPU_SAMPLE_ITERS = (3, 6)


def cpu_reference_step(size: int, workers: int, replicas: int = 1):
    """One bounded CPU sample of `replicas` steps (one per GPU of the other arm): every method at 3 and at 6 iterations, the jobs
    spread over `workers` processes (the reference is single-threaded per pair).  Returns the measured sample and the time the
    full step (the GPU arm's iteration caps) takes at the measured set-up and per-iteration costs."""
    import multiprocessing as mp
    jobs = [(m, size, r) for r in range(replicas) for m in METHODS]
    t0 = time.perf_counter()
    if workers > 1:
        with mp.get_context("fork").Pool(workers) as pool:
            res = pool.map(_cpu_one, jobs, chunksize=1)
    else:
        res = [_cpu_one(j) for j in jobs]
    wall = time.perf_counter() - t0
    n = size * size
    per, full_times = {}, []
    for method, its, ts, kind in res:
        per_iter = max((ts[1] - ts[0]) / max(its[1] - its[0], 1), 1e-9)
        setup = max(ts[0] - its[0] * per_iter, 0.0)
        full = setup + NITER[method] * per_iter
        full_times.append(full)
        d = per.setdefault(method, {"setup_s": 0.0, "per_iteration_s": 0.0, "full_step_s": 0.0, "sample_s": 0.0, "sample_iterations": 0, "n": 0})
        d["setup_s"] += setup; d["per_iteration_s"] += per_iter; d["full_step_s"] += full; d["sample_s"] += sum(ts); d["sample_iterations"] += sum(its); d["n"] += 1
    for d in per.values():
        k = d.pop("n")
        for key in ("setup_s", "per_iteration_s", "full_step_s"):
            d[key] /= k
    # wall time of the full step on `workers` processes: all jobs concurrent when they fit, else bounded below by the total work
    full_wall = max(max(full_times), sum(full_times) / workers)
    work_full = replicas * sum(NITER[m] for m in METHODS) * n
    return {"wall_s": wall, "sample_pixel_iters": sum(sum(r[1]) for r in res) * n, "per_method": per, "kind": res[0][3],
            "full_step_wall_s": full_wall, "full_step_pixel_iters": work_full}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ncpu = os.cpu_count() or 1
    replicas = max(1, args.gpus)                                  # the other arm runs one step per GPU
    workers = max(1, min(len(METHODS) * replicas, ncpu))
    for _ in range(args.warmup if args.warmup < 2 else 1):        # one warm-up pass is enough to page the library in
        cpu_reference_step(min(args.size, 512), workers, 1)
    tot_full, tot_work, tot_wall, tot_sample, last = 0.0, 0, 0.0, 0, None
    for _ in range(args.steps):
        last = cpu_reference_step(args.size, workers, replicas)
        tot_full += last["full_step_wall_s"]; tot_work += last["full_step_pixel_iters"]
        tot_wall += last["wall_s"]; tot_sample += last["sample_pixel_iters"]
    value = tot_work / tot_full / 1e6
    sample = (f"per step: each of the 6 methods at {CPU_SAMPLE_ITERS[0]} and {CPU_SAMPLE_ITERS[1]} iterations at {args.size}^2 ({replicas} replica(s), {workers} processes, "
              f"{tot_wall / args.steps:.1f} s wall) -> set-up and per-iteration cost per method; value = work of the full step at the GPU arm's caps "
              f"{NITER} / (set-up + cap x per-iteration), longest process")
    line = {
        "impl": "reference", "metric": "Mpixel*iter/s (6 registration methods, aggregate)", "value": value, "unit": "Mpixel*iter/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_full / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.size),
        "cpu_baseline": {"value": value, "unit": "Mpixel*iter/s", "cores": workers, "kind": "reference" if last["kind"] == "ref" else "port",
                         "sample": sample, "measured_sample_mpix_iter_s": tot_sample / tot_wall / 1e6, "measured_sample_wall_s_per_step": tot_wall / args.steps,
                         "per_method": last["per_method"], "curvature_dct": "stand-in O(N log N) DCT (fftw3 not installed offline)"},
        "e2e": {"value": value, "unit": "Mpixel*iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "methods": {m: {"mpix_iter_s": args.size * args.size * NITER[m] / v["full_step_s"] / 1e6} for m, v in last["per_method"].items()},
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(size: int) -> dict:
    return {"workload": f"c4_all_methods_{size}x{size}_f32", "size": [size, size], "nscales": 0, "nrefine": 1,
            "niter_cap": NITER, "regparams": PARAMS, "input": "lattice sigma_b=8 (fluid: 6) + texture, shift (1.5,-0.75) px",
            "l2": "six sessions (about 1 GB of fields) are cycled every step, so each method starts from HBM; no extra flush",
            "arithmetic": "level 2 (relaxed engine) unless OF2D_MATH says otherwise; parity against the compiled reference is reported in `parity`"}


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
# algorithmic bytes per pixel per launch of the engine kernels, fp32 (DESIGN.md "Kernels")
KERNEL_BYTES_PER_PX = {
    "hs_iter": 28, "hs_pair": 28, "conv": 16, "conv_logger": 24, "conv_maxabs": 16, "demons_force": 24, "compose": 24, "square": 16,
    "force_conv": 24, "force_conv_maxabs": 24, "compose_conv_logger": 24,
    "sor_tile_elastic": 28, "sor_tile_fluid": 44, "fluid_integrate": 24,
    "curv_rows_fwd": 36, "curv_cols": 32, "curv_rows_inv": 32, "curv_rows_inv_fwd": 60, "regrid_compose": 24, "regrid_rewarp": 36, "final_compose": 24,
}
# SURVEY.md 8(d) budgets the Fluid sweep (its kernel K-a) at 36 B/px: the 8 more are the increment this build materialises
KERNEL_BYTES_PER_PX_SURVEY = {"sor_tile_fluid": 36}
# launches whose event time is not an HBM measurement: a single L2-resident closing launch / self-gated launches that are mostly empty
KERNEL_NOTES = {
    "hs_iter": "one closing launch per refine pass on L2-resident fields (the iterations run in hs_pair): not an HBM measurement",
    "square": "enqueued nsq_cap times per iteration and self-gated: most launches return at once, the average mixes empty and real launches",
    "regrid_compose": "self-gated: runs only after iterations that ask for a regrid", "regrid_rewarp": "self-gated: runs only after iterations that ask for a regrid",
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures (profiles/), 2048^2 fp32
NCU_TRAFFIC_BYTES = {}
try:
    NCU_TRAFFIC_BYTES = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
except (OSError, ValueError):
    pass

BATCH_METHODS = ["thirion", "fluid"]
BATCH_PX = 512
# pairs resident in the engine at a time (measured, 1 GPU, 512 pairs per call, resident / e2e pairs/s -- Thirion: 64: 1617 / 1418, 128: 1665 / 1499,
# 148: 1670 / 1445, 222: 1498 / 1313, 256: 1159 / 1060 (a pair's CTAs times the wave must fill, not exceed, the 444 resident slots of the fused kernels);
# Fluid, whose pairs stop after different iteration counts so that a wave lasts as long as its slowest pair: 32: 1493 / 1410, 64: 1404 / 1328, 128: 1382 / 1298)
BATCH_WAVE = {"thirion": 128, "fluid": 32}
if os.environ.get("OF2D_BENCH_WAVE"):
    BATCH_WAVE = {m: int(os.environ["OF2D_BENCH_WAVE"]) for m in BATCH_WAVE}
BATCH_NITER = {"thirion": 100, "fluid": 60}   # Fluid: the cap at which all sampled pairs are pinned to 1e-3 px (tests/test_configs_gpu.py)


def make_batch_inputs(lo: int, hi: int, size: int):
    from opticalflow2d_b200 import synthetic as S
    n = hi - lo
    R = np.empty((n, size, size)); T = np.empty((n, size, size))
    for k in range(n):
        R[k], T[k] = S.batch_pair(lo + k, size, size)
    return R, T


def _cpu_batch_one(args):
    method, k, size, niter = args
    from oracle import refapi
    from opticalflow2d_b200 import synthetic as S
    kind = "ref" if refapi.available("ref", 32) else "oracle"
    lib = refapi.get(kind, 32)
    R, T = S.batch_pair(k, size, size)
    t0 = time.perf_counter()
    out = lib.register(R, T, REG[method], PARAMS[method], [niter], nscales=0, nrefine=1, verbose=1)
    return len(out["err"]), time.perf_counter() - t0


def cpu_batch_sample(method: str, size: int, niter: int, workers: int):
    import multiprocessing as mp
    jobs = [(method, k, size, niter) for k in range(workers)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(workers) as pool:
        res = pool.map(_cpu_batch_one, jobs)
    wall = time.perf_counter() - t0
    return {"pairs_per_s": workers / wall, "cores": workers, "mean_iterations": float(np.mean([r[0] for r in res])),
            "sample": f"{workers} pairs of {size}^2 ({method}, niter cap {niter}), one process per pair, {wall:.1f} s wall"}


def parity_check(of, sessions, size):
    """The state the timed steps left behind (every session holds the result of a full step) against the fixtures sampled from the
    COMPILED REFERENCE at this size and these caps (tests/golden/full2048_<method>.npz, tests/golden/make_golden_full.py):
    iteration count, regrid trace, max |du| on the 64 x 64 sample lattice.  `ok` = north-star bar (1e-3 px, identical control flow)."""
    out = {"ok": True, "methods": {}}
    for m, s in sessions.items():
        path = os.path.join(ROOT, "tests", "golden", f"full{size}_{m}.npz")
        if not os.path.exists(path):
            out["methods"][m] = {"fixture": None}
            continue
        g = np.load(path)
        if int(g["niter"]) != NITER[m]:
            out["methods"][m] = {"fixture": "cap differs"}
            continue
        tr = s.trace()["levels"][0]
        mo = s.motion()
        st, off = int(g["stride"]), int(g["offset"])
        d = np.abs(mo[off::st, off::st] - g["sample"].astype(np.float64)).max(axis=2)
        pos = off + st * np.arange(d.shape[0])
        inner = (pos >= 8) & (pos < size - 8)     # within 8 px of the border the reference's own fp32 / fp64 builds differ by up to O(1) px (tests/test_fullsize_gpu.py)
        du, du_border = float(d[np.ix_(inner, inner)].max()), float(d.max())
        same_it = tr["iterations"] == len(g["err"])
        same_rg = bool(np.array_equal(np.asarray(tr["regrid_iter"], dtype=int), g["regrid_iter"].astype(int)))
        ok = same_it and same_rg and du <= 1e-3
        out["methods"][m] = {"iterations_match": bool(same_it), "regrid_trace_matches": same_rg, "regrids": int(len(g["regrid_iter"])), "max_du_sampled_px": du,
                             "max_du_sampled_px_incl_border_zone": du_border, "ok": bool(ok)}
        out["ok"] = out["ok"] and bool(ok)
    out["against"] = "tests/golden/full%d_*.npz (compiled reference, fp32, same caps)" % size
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    os.environ.setdefault("OF2D_DEVICE", str(local))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import opticalflow2d_b200 as of
    size = args.size
    n = size * size
    of.set_stream(torch.cuda.current_stream().cuda_stream, 32)
    if of.get_math(32) >= 2:
        # relaxed engine, fp32 fields: the Curvature transforms and the spectrum between the passes are single precision
        # (dct_reg.cuh, namespace rgf) -- SURVEY 8(d)'s 60 B/px layout; the exact engine keeps the reference's double spectrum (92)
        BYTES_PER_PX_ITER["curvature"] = 60
        KERNEL_BYTES_PER_PX.update({"curv_rows_fwd": 28, "curv_cols": 16, "curv_rows_inv": 24, "curv_rows_inv_fwd": 44})

    # host inputs (pinned doubles, as the MEX boundary delivers them) and resident sessions
    sessions, pinned = {}, {}
    for m in METHODS:
        R, T = make_inputs(m, size)
        pr = torch.from_numpy(R).pin_memory(); pt = torch.from_numpy(T).pin_memory()
        pout = torch.empty(2 * n, dtype=torch.float64).pin_memory()
        pinned[m] = (pr, pt, pout)
        s = of.Session((size, size), [NITER[m]], 0, REG[m], PARAMS[m], nrefine=1, verbose=0, bits=32)
        s.set_images_raw(pr.data_ptr(), pt.data_ptr())
        sessions[m] = s
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(e2e: bool, per_method=None):
        iters = {}
        for m in METHODS:
            s = sessions[m]
            pr, pt, pout = pinned[m]
            s.reset()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            if e2e:
                s.set_images_raw(pr.data_ptr(), pt.data_ptr())
            s.estimate()
            if e2e:
                s.motion_raw(pout.data_ptr())
            e1.record()
            iters[m] = s.trace()["total_iterations"]
            if per_method is not None:
                per_method.setdefault(m, []).append((e0, e1))
        return iters

    def timed(e2e: bool):
        for _ in range(args.warmup):
            step(e2e)
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        per = {}
        l0 = of.launch_count(32)
        t0 = time.perf_counter()
        iters = None
        for _ in range(args.steps):
            iters = step(e2e, per)
        barrier()
        wall = time.perf_counter() - t0
        launches = of.launch_count(32) - l0
        clocks = sampler.stop()
        ms = {m: sum(a.elapsed_time(b) for a, b in per[m]) / args.steps for m in METHODS}
        return iters, ms, wall, launches, clocks

    iters, ms, wall, launches, clocks = timed(False)
    parity = parity_check(of, sessions, size) if (rank == 0 and not args.quick) else None
    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "ms": ms, "iterations": iters, "gpu_launches": int(launches)}), flush=True)
        for s in sessions.values():
            s.close()
        return 0
    iters_e, ms_e, wall_e, _, _ = timed(True)

    # e2e, pipelined: the same six registrations through ONE C-ABI call (of2d_sessions_register, include/of2d_host.h): the host ->
    # device copies of method k+1 and the device -> host copy of method k-1 run under the solve of method k (three streams).
    def step_pipelined():
        ss = [sessions[m] for m in METHODS]
        for s_ in ss:
            s_.reset()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        of.Session.register_many_raw(ss, [pinned[m][0].data_ptr() for m in METHODS], [pinned[m][1].data_ptr() for m in METHODS], [pinned[m][2].data_ptr() for m in METHODS])
        e1.record()   # the call returns when the last device -> host copy has landed
        return e0, e1
    for _ in range(args.warmup):
        step_pipelined()
    barrier()
    evp = [step_pipelined() for _ in range(args.steps)]
    barrier()
    te_pipe = sum(a.elapsed_time(b) for a, b in evp) / args.steps * 1e-3
    iters_p = {m: sessions[m].trace()["total_iterations"] for m in METHODS}

    # roofline leg: one extra step with per-kernel CUDA events on the launching stream (not part of the timed steps)
    kern, dom = {}, None
    if rank == 0:
        of.profile_enable(True, 32)
        step(False)
        prof = of.profile_read(32)
        of.profile_enable(False, 32)
        tot_ms = sum(v[1] for v in prof.values()) or 1e-12
        for name, (cnt, tms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
            bpp = KERNEL_BYTES_PER_PX.get(name)
            avg = tms / max(cnt, 1)
            kern[name] = {"launches": cnt, "avg_ms": avg, "share_of_step": tms / tot_ms, "bytes_per_px": bpp,
                          "gbs": (bpp * n / (avg * 1e-3) / 1e9) if (bpp and name not in KERNEL_NOTES) else None}
            if name in KERNEL_BYTES_PER_PX_SURVEY:
                kern[name]["bytes_per_px_survey"] = KERNEL_BYTES_PER_PX_SURVEY[name]
                kern[name]["gbs_survey_bytes"] = KERNEL_BYTES_PER_PX_SURVEY[name] * n / (avg * 1e-3) / 1e9
            if name in KERNEL_NOTES:
                kern[name]["note"] = KERNEL_NOTES[name]
        dom = next((k for k in kern if kern[k]["gbs"]), None)
    for s in sessions.values():
        s.close()
    sessions.clear()
    pinned.clear()

    # ---- fp64 mode (BASELINE.json configs[3] asks for fp64 and fp32): every method once, same inputs and caps
    f64 = {}
    if args.fp64 and rank == 0:
        of.set_stream(torch.cuda.current_stream().cuda_stream, 64)
        for m in METHODS:
            R, T = make_inputs(m, size)
            with of.Session((size, size), [NITER[m]], 0, REG[m], PARAMS[m], nrefine=1, verbose=0, bits=64) as s:
                s.set_images(R, T)
                ts = []
                for rep in range(2):     # first pass warms up (engine creation, caches); second is timed
                    s.reset()
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(); s.estimate(); e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                it = s.trace()["total_iterations"]
            mp = n * it / (ts[-1] * 1e-3) / 1e6
            f64[m] = {"iterations": it, "ms": ts[-1], "mpix_iter_s": mp, "gbs_algorithmic": mp * 1e6 * 2 * BYTES_PER_PX_ITER[m] / 1e9}

    # ---- batched slice registration (BASELINE.json configs[4]): --batch pairs of 512^2 IN TOTAL, sharded over the ranks in
    # contiguous ranges (no collective in the solve).  A rank keeps at most 512 distinct pairs in pinned host memory and goes
    # through its shard in chunks of that size (the inputs of chunk c are pairs lo + c*512 ... modulo the distinct ones).
    batch_res = {}
    if args.batch > 0:
        bp = BATCH_PX
        lo, hi = of.shard_pairs(args.batch, world, rank)
        mine = hi - lo
        B = max(1, min(mine, 512))                      # pairs per call (and distinct inputs held by this rank)
        ncalls = max(1, -(-mine // B))
        Rb, Tb = make_batch_inputs(lo, lo + B, bp)
        prb = torch.from_numpy(Rb).pin_memory(); ptb = torch.from_numpy(Tb).pin_memory()
        pob = torch.empty((B, 2, bp, bp), dtype=torch.float64).pin_memory()
        del Rb, Tb
        for m in BATCH_METHODS:
            bt = of.Batch((bp, bp), B, BATCH_NITER[m], REG[m], PARAMS[m], nrefine=1, wave=min(B, BATCH_WAVE[m]), bits=32)
            bt.set_images_raw(prb.data_ptr(), ptb.data_ptr())
            res = {}
            for leg in ("resident", "e2e"):
                def bstep():
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(ncalls):
                        if leg == "e2e":
                            bt.register_raw(prb.data_ptr(), ptb.data_ptr(), pob.data_ptr())   # streamed: copies of neighbouring waves under the solve
                        else:
                            bt.estimate()
                    e1.record()
                    return e0, e1
                bt.estimate() if leg == "resident" else bt.register_raw(prb.data_ptr(), ptb.data_ptr(), pob.data_ptr())   # warm-up of the leg
                barrier()
                ev = bstep()
                barrier()
                sec = ev[0].elapsed_time(ev[1]) * 1e-3
                its, rg = bt.iterations()
                v = torch.tensor([sec], device="cuda", dtype=torch.float64)
                w = torch.tensor([float(B * ncalls), float(its.sum()) * ncalls], device="cuda", dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(v, op=dist.ReduceOp.MAX)
                    dist.all_reduce(w, op=dist.ReduceOp.SUM)
                res[leg] = {"pairs_per_s": float(w[0]) / float(v[0]), "seconds": float(v[0]),
                            "mpix_iter_s": float(w[1]) * bp * bp / float(v[0]) / 1e6}
                res["mean_iterations"] = float(w[1]) / float(w[0])
                res["pairs"] = int(w[0])
            bt.close()
            batch_res[m] = res
        batch_res["config"] = {"workload": f"c5_batch_{bp}x{bp}", "total_pairs": args.batch, "n_gpus": world, "pairs_per_call": B, "calls_per_rank": ncalls,
                               "niter_cap": BATCH_NITER, "wave": {m: min(B, BATCH_WAVE[m]) for m in BATCH_METHODS}, "h2d_bytes_per_pair": 2 * bp * bp * 8, "d2h_bytes_per_pair": 2 * bp * bp * 8,
                               "e2e": "of2d_batch_register: pinned host doubles in, planar doubles out, copies of wave k+1 / k-1 under the solve of wave k",
                               "sharding": "contiguous pair ranges per rank, no collective in the solve"}
        del prb, ptb, pob

    def agg(ms_map, it_map):
        t = sum(ms_map.values()) * 1e-3
        px = sum(it_map[m] for m in METHODS) * n
        return px, t

    px, t = agg(ms, iters)
    pxe, te = agg(ms_e, iters_e)
    pxp = sum(iters_p[m] for m in METHODS) * n
    # max over ranks of the step time, sum over ranks of the work
    if world > 1:
        v = torch.tensor([t, te, te_pipe], device="cuda", dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        t, te, te_pipe = float(v[0]), float(v[1]), float(v[2])
        w = torch.tensor([px, pxe, pxp], device="cuda", dtype=torch.float64)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        px, pxe, pxp = float(w[0]), float(w[1]), float(w[2])

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        methods = {}
        for m in METHODS:
            mp = n * iters[m] / (ms[m] * 1e-3) / 1e6
            gbs = mp * 1e6 * BYTES_PER_PX_ITER[m] / 1e9
            methods[m] = {"iterations": iters[m], "ms": ms[m], "mpix_iter_s": mp, "gbs_algorithmic": gbs, "frac_of_hbm_peak": gbs / peak,
                          "bytes_per_px_iter": BYTES_PER_PX_ITER[m], "bytes_per_px_iter_survey": BYTES_PER_PX_ITER_SURVEY[m],
                          "frac_of_hbm_peak_survey_bytes": mp * 1e6 * BYTES_PER_PX_ITER_SURVEY[m] / 1e9 / peak, "e2e_ms": ms_e[m]}
        ncpu = os.cpu_count() or 1
        workers = max(1, min(len(METHODS), ncpu))
        c = cpu_reference_step(size, workers, 1)
        cpu = {"value": c["full_step_pixel_iters"] / c["full_step_wall_s"] / 1e6, "unit": "Mpixel*iter/s", "cores": workers,
               "kind": "reference" if c["kind"] == "ref" else "port",
               "sample": (f"each of the 6 methods at {CPU_SAMPLE_ITERS[0]} and {CPU_SAMPLE_ITERS[1]} iterations at {size}^2, {workers} processes, {c['wall_s']:.1f} s wall -> set-up and "
                          f"per-iteration cost per method; value = work of the full step at the GPU arm's caps / (set-up + cap x per-iteration), longest process"),
               "measured_sample_mpix_iter_s": c["sample_pixel_iters"] / c["wall_s"] / 1e6,
               "per_method": c["per_method"],
               "per_method_mpix_iter_s": {m: n * NITER[m] / v["full_step_s"] / 1e6 for m, v in c["per_method"].items()},
               "curvature_dct": "stand-in O(N log N) DCT (fftw3 not installed offline)"}
        if args.batch > 0:
            bw = max(1, min(ncpu, 16))
            cpu["batch"] = {m: cpu_batch_sample(m, BATCH_PX, BATCH_NITER[m], bw) for m in BATCH_METHODS}
        line = {
            "metric": "Mpixel*iter/s (6 registration methods, aggregate)", "value": px / t / 1e6, "unit": "Mpixel*iter/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(size),
            "e2e": {"value": pxp / te_pipe / 1e6, "unit": "Mpixel*iter/s", "h2d_bytes_per_step": len(METHODS) * 2 * n * 8,
                    "d2h_bytes_per_step": len(METHODS) * 2 * n * 8, "ms_per_step": 1e3 * te_pipe,
                    "how": "of2d_sessions_register: the six registrations in one C-ABI call, pinned host doubles in / planar doubles out, the copies of the neighbouring methods under each solve",
                    "sequential": {"value": pxe / te / 1e6, "ms_per_step": 1e3 * te, "how": "set_images, estimate, get_motion per method, one after the other (every copy exposed)"}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s",
                         "frac": kern[dom]["gbs"] / peak, "traffic": NCU_TRAFFIC_BYTES.get(dom), "peak_source": peak_src,
                         "frac_survey_bytes": (kern[dom].get("gbs_survey_bytes") or kern[dom]["gbs"]) / peak,
                         "share_of_step": kern[dom]["share_of_step"], "avg_launch_ms": kern[dom]["avg_ms"],
                         "algorithmic_bytes_per_launch": kern[dom]["bytes_per_px"] * n,
                         "how": "CUDA events around every engine launch of one extra (untimed) step, on the launching stream"} if dom else None,
            "kernels": kern,
            "methods": methods,
            "methods_f64": {m: dict(v, frac_of_hbm_peak=v["gbs_algorithmic"] / peak) for m, v in f64.items()},
            "parity": parity,
            "math_level": {0: "strict", 1: "exact", 2: "relaxed"}.get(of.get_math(32)),
            "cpu_baseline": cpu,
            "wall_s_timed_region": wall,
            "loaded_libraries": [os.path.relpath(p, ROOT) for p in of.loaded_libraries()],
            # LAST key (the driver keeps the tail of the line): BASELINE.json configs[4], pairs/s over all ranks
            "batch": batch_res,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=4096, help="pairs of 512^2 in the batched leg IN TOTAL, sharded over the ranks (0 = skip); BASELINE configs[4]: 4096")
    ap.add_argument("--no-fp64", dest="fp64", action="store_false", help="skip the fp64-mode leg")
    ap.add_argument("--quick", action="store_true", help="timed region only (no e2e leg, kernel microbench or CPU baseline): for ncu launch lists")
    ap.add_argument("--methods", default=",".join(list(METHODS)), help="comma-separated subset (profiling only; the default is the benchmark)")
    args = ap.parse_args()
    METHODS[:] = [m for m in args.methods.split(",") if m]
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
