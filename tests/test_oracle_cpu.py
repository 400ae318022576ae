"""CPU suite (-m "not gpu"): pins the ORACLE.  The C restatement (oracle/of2d_oracle.c) must reproduce
the golden vectors generated from the compiled reference bit for bit, and -- when the compiled
reference itself is available (oracle/_ref, built here from /root/reference) -- match it on fresh
seeded inputs for every primitive.  The fftw stand-in is pinned against scipy's DCT-II/III."""
import glob
import os

import numpy as np
import pytest
import scipy.fft

from opticalflow2d_b200 import synthetic as S
from oracle import refapi

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith(("full2048", "c2_", "c3_", "c5_", "c5f64_", "demo_")))   # sampled 2048^2 fixtures: tests/test_fullsize_gpu.py
NP = {32: np.float32, 64: np.float64}


@pytest.fixture(scope="module", autouse=True)
def _build_oracle():
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-s", "-C", os.path.join(root, "oracle"), "oracle"], check=True)


def test_golden_fixtures_present():
    assert len(GOLDEN) >= 9


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
@pytest.mark.parametrize("bits", [32, 64])
def test_oracle_reproduces_reference_golden_vectors(path, bits):
    g = np.load(path)
    orc = refapi.get("oracle", bits)
    out = orc.register(g["Iref"], g["Imov"], int(g["reg"]), list(g["regparams"]), list(g["niter"]), nscales=int(g["nscales"]),
                       nrefine=int(g["nrefine"]), verbose=1)
    assert np.array_equal(out["motion"], g[f"motion_{bits}"])
    assert np.array_equal(out["warped"], g[f"warped_{bits}"])
    assert np.array_equal(out["err"], g[f"err_{bits}"])
    assert np.array_equal(out["err_iter"], g[f"err_iter_{bits}"])
    assert np.array_equal(out["regrid_iter"], g[f"regrid_iter_{bits}"])
    assert np.array_equal(out["regrid_minjac"], g[f"regrid_minjac_{bits}"])
    assert np.array_equal(out["fluid_dt"], g[f"fluid_dt_{bits}"])


@pytest.mark.parametrize("n", [2, 8, 64, 2048, 12, 100, 7])
@pytest.mark.parametrize("kind", [2, 3])
def test_dct_standin_matches_fftw_definitions(n, kind):
    """REDFT10 / REDFT01 as published in the FFTW manual == scipy.fft.dct(type=2|3, norm=None)."""
    rng = np.random.default_rng(n + kind)
    x = rng.standard_normal(n)
    got = refapi.get("oracle", 32).dct1d(x, kind)
    want = scipy.fft.dct(x, type=kind, norm=None)
    assert np.max(np.abs(got - want)) <= 2e-15 * n * np.max(np.abs(want))
    # definition check on a small case, straight from the formula
    if n <= 12:
        j = np.arange(n)
        if kind == 2:
            direct = np.array([2 * np.sum(x * np.cos(np.pi * (j + 0.5) * k / n)) for k in range(n)])
        else:
            direct = np.array([x[0] + 2 * np.sum(x[1:] * np.cos(np.pi * j[1:] * (k + 0.5) / n)) for k in range(n)])
        assert np.allclose(got, direct, rtol=0, atol=1e-12)


def test_dct_roundtrip_scale_is_2n_per_dimension():
    """DCT-III(DCT-II(x)) = 2n x, hence the reference's division by 4*N (OpticalFlowCurvature.cpp:117)."""
    orc = refapi.get("oracle", 64)
    x = np.random.default_rng(0).standard_normal(32)
    assert np.allclose(orc.dct1d(orc.dct1d(x, 2), 3), 64 * x, atol=1e-11)


needs_ref = pytest.mark.skipif(not (refapi.available("ref", 32) and refapi.available("ref", 64)),
                               reason="compiled reference (oracle/_ref) not built")


@needs_ref
@pytest.mark.parametrize("bits", [32, 64])
def test_oracle_primitives_match_compiled_reference(bits):
    ref, orc = refapi.get("ref", bits), refapi.get("oracle", bits)
    dimx, dimy = 53, 38
    R, T = S.make_pair(dimx, dimy, "lattice", smooth=True, sigma_b=5.0)
    R, T = R.astype(NP[bits]), T.astype(NP[bits])
    u = S.random_motion(dimx, dimy, 3.0, 1).astype(NP[bits])
    v = S.random_motion(dimx, dimy, 40.0, 2).astype(NP[bits])   # large: exercises the out-of-range branch
    for a, b in ((u, v), (v, u)):
        assert np.array_equal(ref.warp2d(R, a), orc.warp2d(R, a))
        assert np.array_equal(ref.accumulate(a, b), orc.accumulate(a, b))
    for w, sigma in ((5, 1.5), (3, 0.7), (4, 1.0), (7, 2.5)):
        assert np.array_equal(ref.gaussian_kernel(w, sigma), orc.gaussian_kernel(w, sigma))
        assert np.array_equal(ref.convolute_motion(u, w, sigma), orc.convolute_motion(u, w, sigma))
    for amp in (0.05, 1.0, 7.0):
        assert np.array_equal(ref.exp(u * NP[bits](amp)), orc.exp(u * NP[bits](amp)))
    assert ref.norm_maxabs(u) == orc.norm_maxabs(u)
    jr, mr = ref.jacobian(u)
    jo, mo = orc.jacobian(u)
    assert np.array_equal(jr, jo) and mr == mo
    gr, ir = ref.derivatives(R, T)
    go, io = orc.derivatives(R, T)
    assert np.array_equal(gr, go) and np.array_equal(ir, io)
    assert np.array_equal(ref.set_image(R.astype(np.float64) * np.pi), orc.set_image(R.astype(np.float64) * np.pi))
    assert np.array_equal(ref.copy_motion_to_input(u), orc.copy_motion_to_input(u))
    for up, shp in ((False, (19, 26)), (False, (9, 13)), (True, (80, 110))):
        assert np.array_equal(ref.image_resample(R, shp, up), orc.image_resample(R, shp, up))
        assert np.array_equal(ref.motion_resample(u, shp, up), orc.motion_resample(u, shp, up))
    seq = np.stack([u, v, u * NP[bits](0.5), u * NP[bits](0.5)])
    assert np.array_equal(ref.logger(seq), orc.logger(seq))


@needs_ref
@pytest.mark.parametrize("bits", [32, 64])
@pytest.mark.parametrize("reg,params", [(0, [0.5]), (1, [0.25]), (1, [0.3, 0.8]), (2, [1.0, 0.25]), (2, [0.7, 0.1, 1.1]),
                                        (3, [1, 0.25, 1.5, 1.5, 5, 0]), (3, [1, 0.25, 2.0, 1.0, 3, 1]), (4, [1, 2.0, 1.5, 1.5, 5]),
                                        (5, [0.1, 0.0]), (5, [0.2, 0.05, 0.9])])
def test_oracle_solver_steps_match_compiled_reference(bits, reg, params):
    ref, orc = refapi.get("ref", bits), refapi.get("oracle", bits)
    dimx, dimy = 48, 40
    R, T = S.make_pair(dimx, dimy, "lattice", smooth=True, sigma_b=5.0)
    R, T = R.astype(NP[bits]), T.astype(NP[bits])
    u0 = S.random_motion(dimx, dimy, 0.2, 3).astype(NP[bits])
    assert np.array_equal(ref.solver_steps(reg, params, R, T, u0, 4), orc.solver_steps(reg, params, R, T, u0, 4))


@pytest.mark.parametrize("kind", ["oracle"] + (["ref"] if refapi.available("ref", 32) else []))
def test_error_conventions(kind):
    """wrong nparams -> std::invalid_argument (2); divide by zero -> std::runtime_error (3);
    bad call shape -> mexErrMsgTxt (3)  (SURVEY 8b error conventions)."""
    lib = refapi.get(kind, 32)
    R, T = S.make_pair(16, 16, "blob")
    with pytest.raises(refapi.RefError) as e:
        lib.register(R, T, 0, [0.5, 1.0], [3])
    assert e.value.code == 2
    flat = np.ones((16, 16))
    with pytest.raises(refapi.RefError) as e:
        lib.register(flat, flat, 3, [1, 0.25, 1.5, 1.5, 5, 0], [3])
    assert e.value.code == 3 and "Divide by zero" in e.value.msg
    assert lib.mex_badcall(1, 0) == 3      # motion requested with no live object
    assert lib.mex_badcall(2, 2) == 3


def test_three_iterations_always_run():
    """error[0] is 0 and the break needs iter > 1 (Logger.cpp:39, ImageRegistrationOpticalFlow.cpp:131-134)."""
    orc = refapi.get("oracle", 32)
    R, _ = S.make_pair(24, 24, "blob")
    out = orc.register(R, R, 0, [0.5], [50])       # identical images: zero motion, still three iterations
    assert len(out["err"]) == 3 and np.all(out["motion"] == 0)


@pytest.mark.parametrize("bits", [32, 64])
def test_oracle_surface_functions_match_compiled_reference(bits):
    """Image::sum/max/min/normalize, Motion::Neumann_/Dirichlet_boundaryconditions, Kernel::set_average --
    the restatement against the compiled reference, bit for bit.  Image::convolute is NOT comparable: the reference leaves its
    float accumulator uninitialised (src/Field.tpp:240) and the compiled code carries it from pixel to pixel; the restatement
    (and the CUDA path) start from zero, which this test documents by checking the restatement against the definition."""
    if not refapi.available("ref", bits):
        pytest.skip("oracle/_ref not built")
    ref, orc = refapi.get("ref", bits), refapi.get("oracle", bits)
    rng = np.random.default_rng(11)
    for dimx, dimy in ((52, 37), (37, 52), (5, 9)):
        img = rng.uniform(-0.3, 1.2, (dimy, dimx))
        u = rng.normal(size=(dimy, dimx, 2))
        assert ref.image_stats(img) == orc.image_stats(img)
        assert np.array_equal(ref.image_normalize(img), orc.image_normalize(img))
        for kind in (0, 1):
            assert np.array_equal(ref.boundary_conditions(u, kind), orc.boundary_conditions(u, kind))
    for w in (3, 4, 5, 7):
        assert np.array_equal(ref.average_kernel(w), orc.average_kernel(w))
    # Image::convolute, zero-initialised: interior pixels of a constant image stay constant, a delta spreads the kernel
    img = np.full((20, 24), 0.75)
    out = orc.convolute_image(img, 5, 1.5)
    assert np.allclose(out[3:-3, 3:-3], 0.75, rtol=1e-6)
    delta = np.zeros((21, 21)); delta[10, 10] = 1.0
    k = orc.gaussian_kernel(5, 1.5)
    assert np.allclose(orc.convolute_image(delta, 5, 1.5)[8:13, 8:13], k[::-1, ::-1], rtol=1e-6, atol=1e-9)
