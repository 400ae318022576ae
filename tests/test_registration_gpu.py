"""Full-schedule parity (-m gpu): the 5-call MEX protocol of test_opticalflow2d.m:42-59 replayed
through this repo's mexFunction, against (a) the golden vectors generated from the compiled
reference and (b) the C oracle on fresh inputs, for all six methods, fp32 and fp64.

Bars (BASELINE.json north_star): fp32 max |du| <= 1e-3 px and final SSD relative error <= 1e-4;
fp64 max |du| <= 1e-6 px; identical control flow (iteration counts, regrid iterations).  Strict mode
is expected to do far better (bit-exact unless a reduction-driven decision differs); both the
default fast mode and strict mode are held to the north-star bars."""
import glob
import os

import numpy as np
import pytest

import opticalflow2d_b200 as of
from gpu_common import maxdiff, oracle
from opticalflow2d_b200 import synthetic as S

pytestmark = pytest.mark.gpu
GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith(("full2048", "c2_", "c3_", "c5_", "c5f64_", "demo_")))   # sampled 2048^2 fixtures: tests/test_fullsize_gpu.py
TOL_PX = {32: 1e-3, 64: 1e-6}


def ssd(a, b):
    return float(np.sum((np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)) ** 2))


def run_mex(bits, R, T, reg, params, niter, nscales=0, nrefine=1, strict=True):
    of.set_strict(strict, bits)
    f = of.OpticalFlow2d(bits)
    dimy, dimx = R.shape
    f.init((dimx, dimy), niter, nscales, reg, params, nrefine, verbose=0)
    try:
        f.register(R, T)
        motion = f.motion()
        warped = f.warp(T)
        trace = f.trace()
    finally:
        f.close()
    return motion, warped, trace


def check_against(bits, R, motion, warped, trace, want_motion, want_warped, want_err, want_regrid):
    iters = [l["iterations"] for l in trace["levels"]]
    assert sum(iters) == len(want_err), (iters, len(want_err))
    got_regrid = np.concatenate([l["regrid_iter"] for l in trace["levels"]]) if trace["levels"] else np.zeros(0)
    assert np.array_equal(got_regrid, want_regrid)
    got_err = np.concatenate([l["err"] for l in trace["levels"]])
    assert np.allclose(got_err, want_err, rtol=5e-4, atol=1e-9)
    assert maxdiff(motion, want_motion) <= TOL_PX[bits]
    s_ref, s_got = ssd(want_warped, R), ssd(warped, R)
    assert abs(s_got - s_ref) <= 1e-4 * s_ref + 1e-12


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
@pytest.mark.parametrize("bits", [32, 64])
@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fast"])
def test_mex_protocol_matches_reference_golden_vectors(path, bits, strict):
    g = np.load(path)
    motion, warped, trace = run_mex(bits, g["Iref"], g["Imov"], int(g["reg"]), list(g["regparams"]), list(g["niter"]),
                                    int(g["nscales"]), int(g["nrefine"]), strict)
    check_against(bits, g["Iref"], motion, warped, trace, g[f"motion_{bits}"], g[f"warped_{bits}"], g[f"err_{bits}"], g[f"regrid_iter_{bits}"])


CONFIGS = [
    # the BASELINE.json configurations at sizes the CPU oracle finishes in seconds
    ("C1 horn-schunck blob", 256, 256, dict(kind="blob"), 0, [0.5], [200]),
    ("C2 thirion smooth", 192, 160, dict(kind="lattice", sigma_b=8.0, smooth=True), 3, [1, 0.25, 1.5, 1.5, 5, 0], [120]),
    ("C3 diffeomorphic", 160, 128, dict(kind="lattice", sigma_b=8.0), 4, [1, 2.0, 1.5, 1.5, 5], [40]),
    ("C4a curvature", 256, 128, dict(kind="lattice", sigma_b=8.0), 1, [0.25, 1.0], [50]),
    ("C4b elastic", 200, 168, dict(kind="lattice", sigma_b=8.0), 2, [1.0, 0.25], [50]),
    ("C4c fluid", 160, 128, dict(kind="lattice", sigma_b=6.0), 5, [0.1, 0.0], [120]),
]


@pytest.mark.parametrize("name,dimx,dimy,kw,reg,params,niter", CONFIGS, ids=[c[0] for c in CONFIGS])
@pytest.mark.parametrize("bits", [32, 64])
def test_baseline_configs_against_oracle(name, dimx, dimy, kw, reg, params, niter, bits):
    R, T = S.make_pair(dimx, dimy, **kw)
    want = oracle(bits).register(R, T, reg, params, niter, verbose=1)
    motion, warped, trace = run_mex(bits, R, T, reg, params, niter, strict=True)
    check_against(bits, R, motion, warped, trace, want["motion"], want["warped"], want["err"], want["regrid_iter"])
    if reg in (0, 2, 3, 4, 5) and bits == 32:
        # everything but the reductions is order-preserving: strict mode should be exact here
        assert maxdiff(motion, want["motion"]) == 0.0


def test_error_conventions_match_reference():
    f = of.OpticalFlow2d(32)
    with pytest.raises(of.OF2DError) as e:       # wrong nparams -> std::invalid_argument
        f.init((16, 16), [3], 0, 0, [0.5, 1.0])
    assert e.value.code == 2
    with pytest.raises(of.OF2DError) as e:       # no live object -> mexErrMsgTxt
        f.motion_shape = None
        f.call(1, [])
    assert e.value.code == 3 and "invalid number of input and output" in e.value.msg
    # divide by zero: Demons on identical flat images (coord2d.h:95-100 via Demons.cpp:57)
    flat = np.ones((16, 16))
    f.init((16, 16), [3], 0, 3, [1, 0.25, 1.5, 1.5, 5, 0])
    try:
        with pytest.raises(of.OF2DError) as e:
            f.register(flat, flat)
        assert e.value.code == 3 and "Divide by zero" in e.value.msg
    finally:
        f.close()


def test_second_register_call_warm_starts_like_reference():
    """motion[nscales] is not reset between estimate_motion() calls (ImageRegistration.cpp:135-139)."""
    R, T = S.make_pair(64, 64, "blob")
    f = of.OpticalFlow2d(32)
    of.set_strict(True)
    f.init((64, 64), [15], 0, 0, [0.5])
    try:
        f.register(R, T)
        m1 = f.motion()
        f.register(R, T)
        m2 = f.motion()
    finally:
        f.close()
    assert maxdiff(m1, m2) > 1e-4


@pytest.mark.parametrize("bits", [32, 64])
def test_sessions_register_equals_the_three_separate_calls(bits):
    """of2d_sessions_register (include/of2d_host.h): set_images + estimate + get_motion of several sessions in one call with the
    copies of the neighbouring jobs under each solve -- same motion, same traces as the separate calls, for different methods,
    sizes and pyramid depths in one list."""
    jobs = [((96, 64), [6], 0, of.DIFFUSION, [0.5]), ((128, 96), [5, 6], 1, of.THIRION, [1.0, 0.25, 1.5, 1.5, 5, 0]), ((96, 64), [8], 0, of.FLUID, [0.1, 0.0]),
            ((64, 64), [4], 0, of.CURVATURE, [0.25, 1.0]), ((160, 96), [4, 4, 5], 2, of.ELASTIC, [1.0, 0.25])]
    pairs = [S.make_pair(d[0], d[1], "lattice", shift=(1.0 + 0.2 * k, -0.5), smooth=True) for k, (d, *_r) in enumerate(jobs)]
    want, want_it = [], []
    for (dims, niter, nscales, reg, params), (R, T) in zip(jobs, pairs):
        with of.Session(dims, niter, nscales, reg, params, bits=bits) as s:
            s.set_images(R, T)
            s.estimate()
            want.append(s.motion()); want_it.append(s.trace()["total_iterations"])
    sessions = [of.Session(dims, niter, nscales, reg, params, bits=bits) for dims, niter, nscales, reg, params in jobs]
    try:
        Rs = [np.ascontiguousarray(R, dtype=np.float64) for R, _ in pairs]
        Ts = [np.ascontiguousarray(T, dtype=np.float64) for _, T in pairs]
        outs = [np.zeros((2,) + R.shape) for R in Rs]
        for rep in range(2):   # the second call reuses the staging buffers and the side streams
            for s in sessions:
                s.reset()
            of.Session.register_many_raw(sessions, [a.ctypes.data for a in Rs], [a.ctypes.data for a in Ts], [o.ctypes.data for o in outs])
            for k, s in enumerate(sessions):
                got = np.stack([outs[k][0], outs[k][1]], axis=-1)
                assert s.trace()["total_iterations"] == want_it[k]
                assert np.array_equal(got, want[k]), (rep, k, float(np.abs(got - want[k]).max()))
    finally:
        for s in sessions:
            s.close()


def test_sessions_register_rejects_bad_lists():
    """The same session twice in one list (its images would be overwritten while it is being solved) and null buffers are refused
    before anything is enqueued."""
    R, T = S.make_pair(64, 48, "lattice", shift=(1.0, 0.5), smooth=True)
    Rd, Td = np.ascontiguousarray(R, dtype=np.float64), np.ascontiguousarray(T, dtype=np.float64)
    out = np.zeros((2,) + R.shape)
    with of.Session((64, 48), [4], 0, of.DIFFUSION, [0.5]) as s:
        with pytest.raises(of.OF2DError):
            of.Session.register_many_raw([s, s], [Rd.ctypes.data] * 2, [Td.ctypes.data] * 2, [out.ctypes.data] * 2)
        with pytest.raises(of.OF2DError):
            of.Session.register_many_raw([s], [0], [Td.ctypes.data], [out.ctypes.data])
        of.Session.register_many_raw([s], [Rd.ctypes.data], [Td.ctypes.data], [out.ctypes.data])   # the session is still usable
        assert s.trace()["total_iterations"] == 4
