"""Device-resident iteration engine (-m gpu): fast mode (engine) against strict mode (the reference's loop
replayed step by step, bit-exact) and against the CPU oracle.

 * Diffusion / Thirion / Diffeomorphic: the engine's field arithmetic is the reference's, unfused and in
   order, so the motion must be bit-identical to strict mode.
 * Curvature: same mathematics, different FFT butterfly order -> <= 1e-6 px (fp32) / 1e-10 px (fp64).
 * Elastic / Fluid: overlapped-tile sweep (csrc/sor_tile.cuh) -> within the north-star tolerance of the
   exact wavefront (1e-3 px fp32, 1e-6 px fp64), in practice ~1e-7 / 1e-13.
 * identical control flow: iteration counts, regrid iterations, Logger error series.
 * parameter sets whose sweep does not contract (omega close to 2) fall back to the exact path."""
import numpy as np
import pytest

import opticalflow2d_b200 as of
from gpu_common import maxdiff, oracle
from opticalflow2d_b200 import synthetic as S

# these tests pin the EXACT engine (arithmetic level 1) bit for bit; the default relaxed engine: tests/test_relaxed_gpu.py
pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("exact_engine")]


def run(bits, strict, dims, R, T, reg, params, niter, nscales=0, nrefine=1):
    of.set_strict(strict, bits)
    with of.Session(dims, niter, nscales, reg, params, nrefine=nrefine, verbose=0, bits=bits) as s:
        s.set_images(R, T)
        s.estimate()
        return s.motion(), s.trace()


def series(trace, key):
    return np.concatenate([np.asarray(l[key], dtype=np.float64) for l in trace["levels"]]) if trace["levels"] else np.zeros(0)


EXACT = [
    (of.DIFFUSION, [0.5], [25]),
    (of.THIRION, [1.0, 0.25, 1.5, 1.5, 5, 0], [20]),
    (of.THIRION, [1.0, 0.25, 1.5, 1.5, 5, 1], [12]),          # addition instead of composition
    (of.THIRION, [1.0, 0.5, 2.0, 1.0, 7, 0], [10]),           # 7 x 7 kernels
    (of.THIRION, [1.0, 0.5, 2.0, 1.0, 4, 0], [8]),            # even width: generic convolution path
    (of.DIFFEOMORPHIC, [1.0, 2.0, 1.5, 1.5, 5], [15]),        # sigma_x = 2: squarings are active
]


@pytest.mark.parametrize("bits", [32, 64])
@pytest.mark.parametrize("reg,params,niter", EXACT, ids=[f"{of.METHOD_NAMES[c[0]]}-{k}" for k, c in enumerate(EXACT)])
@pytest.mark.parametrize("dimx,dimy", [(160, 96), (97, 131)])
def test_engine_is_bit_identical_to_strict(bits, reg, params, niter, dimx, dimy):
    R, T = S.make_pair(dimx, dimy, "lattice", shift=(1.5, -0.75), smooth=True)
    mf, tf = run(bits, False, (dimx, dimy), R, T, reg, params, niter)
    ms, ts = run(bits, True, (dimx, dimy), R, T, reg, params, niter)
    assert tf["total_iterations"] == ts["total_iterations"]
    assert np.allclose(series(tf, "err"), series(ts, "err"), rtol=1e-5, atol=1e-12)
    assert maxdiff(mf, ms) == 0.0


# Diffusion runs two Jacobi steps per launch (k_hs_pair, temporal blocking).  These pairs converge and break after an odd
# (145, 147) or an even (188, 192) number of iterations on the CPU reference: the break lands on the first / the second
# step of a two-step launch, and on the first one the launch must leave the state to the single-step kernel that follows.
BREAKS = [("blob", (0.3, -0.2), 0.02, 145), ("lattice", (0.4, 0.1), 0.02, 147), ("lattice", (0.4, 0.1), 0.05, 188), ("blob", (1.0, 0.5), 0.02, 192)]


@pytest.mark.parametrize("kind,shift,alpha,want_iter", BREAKS, ids=[f"{b[0]}-{b[3]}" for b in BREAKS])
def test_two_step_diffusion_breaks_where_the_reference_does(kind, shift, alpha, want_iter):
    dimx, dimy = 64, 48
    R, T = S.make_pair(dimx, dimy, kind, shift=shift)
    mf, tf = run(32, False, (dimx, dimy), R, T, of.DIFFUSION, [alpha], [300])
    want = oracle(32).register(R, T, of.DIFFUSION, [alpha], [300], nscales=0, nrefine=1, verbose=1)
    assert len(want["err"]) == want_iter
    assert tf["total_iterations"] == want_iter
    assert np.allclose(series(tf, "err"), np.asarray(want["err"], dtype=np.float64), rtol=5e-4, atol=1e-12)
    assert maxdiff(mf, want["motion"]) == 0.0


@pytest.mark.parametrize("niter", [1, 2, 3, 8, 9])
def test_two_step_diffusion_iteration_caps(niter):
    dimx, dimy = 97, 70
    R, T = S.make_pair(dimx, dimy, "lattice", shift=(1.5, -0.75))
    mf, tf = run(32, False, (dimx, dimy), R, T, of.DIFFUSION, [0.5], [niter])
    want = oracle(32).register(R, T, of.DIFFUSION, [0.5], [niter], nscales=0, nrefine=1, verbose=1)
    assert tf["total_iterations"] == len(want["err"])
    assert maxdiff(mf, want["motion"]) == 0.0


@pytest.mark.parametrize("bits", [32, 64])
@pytest.mark.parametrize("dimx,dimy", [(128, 128), (256, 64), (512, 512), (1024, 512)])
def test_engine_curvature_fast_dct(bits, dimx, dimy):
    R, T = S.make_pair(dimx, dimy, "lattice", shift=(1.5, -0.75))
    mf, tf = run(bits, False, (dimx, dimy), R, T, of.CURVATURE, [0.25, 1.0], [12])
    ms, ts = run(bits, True, (dimx, dimy), R, T, of.CURVATURE, [0.25, 1.0], [12])
    assert tf["total_iterations"] == ts["total_iterations"]
    assert maxdiff(mf, ms) <= (1e-6 if bits == 32 else 1e-10)


SOR = [
    (of.ELASTIC, [1.0, 0.25], [20]),
    (of.ELASTIC, [1.0, 0.0, 0.9], [10]),
    (of.ELASTIC, [0.5, 2.0, 0.5], [10]),
    (of.ELASTIC, [0.37, 0.13], [12]),          # non-dyadic parameters: the tile coefficients must round like the reference's float expression
    (of.FLUID, [0.37, 0.13], [12]),
    (of.FLUID, [0.1, 0.0], [30]),
    (of.FLUID, [0.2, 0.1, 0.8], [20]),
]


@pytest.mark.parametrize("bits", [32, 64])
@pytest.mark.parametrize("reg,params,niter", SOR, ids=[f"{of.METHOD_NAMES[c[0]]}-{k}" for k, c in enumerate(SOR)])
@pytest.mark.parametrize("dimx,dimy", [(200, 300), (320, 130)])
def test_tiled_sweep_matches_exact_wavefront(bits, reg, params, niter, dimx, dimy):
    R, T = S.make_pair(dimx, dimy, "lattice", shift=(1.5, -0.75), sigma_b=6.0)
    mf, tf = run(bits, False, (dimx, dimy), R, T, reg, params, niter)
    ms, ts = run(bits, True, (dimx, dimy), R, T, reg, params, niter)
    assert tf["total_iterations"] == ts["total_iterations"]
    assert np.array_equal(series(tf, "regrid_iter"), series(ts, "regrid_iter"))
    assert np.allclose(series(tf, "err"), series(ts, "err"), rtol=5e-4, atol=1e-10)
    assert maxdiff(mf, ms) <= (1e-3 if bits == 32 else 1e-6)      # north-star bar
    # what the tiles actually achieve here: Elastic is a contraction (rounding-level differences stay there); Fluid's
    # regridding + explicit Euler steps amplify 1-ulp differences ~10x per 10 iterations
    achieved = {(of.ELASTIC, 32): 2e-5, (of.ELASTIC, 64): 1e-9, (of.FLUID, 32): 5e-4, (of.FLUID, 64): 1e-9}[(reg, bits)]
    assert maxdiff(mf, ms) <= achieved


def test_non_contracting_sweep_falls_back_to_exact_path(capfd):
    dimx, dimy = 96, 80
    R, T = S.make_pair(dimx, dimy, "lattice", shift=(1.0, -0.5))
    params = [1.0, 0.25, 1.9]           # over-relaxation: the tile halos would not converge
    mf, tf = run(32, False, (dimx, dimy), R, T, of.ELASTIC, params, [6])
    ms, ts = run(32, True, (dimx, dimy), R, T, of.ELASTIC, params, [6])
    assert maxdiff(mf, ms) == 0.0
    want = oracle(32).register(R, T, of.ELASTIC, params, [6], nscales=0, nrefine=1, verbose=1)
    assert maxdiff(mf, want["motion"]) == 0.0
    assert "runs on the per-iteration path" in capfd.readouterr().err      # the 150x slower path is taken loudly, not silently


@pytest.mark.parametrize("reg,params,niter", [(of.DIFFUSION, [0.5], [6, 8, 10]), (of.FLUID, [0.1, 0.0], [8, 8, 12]), (of.THIRION, [1.0, 0.25, 1.5, 1.5, 5, 0], [5, 6, 8])],
                         ids=["diffusion", "fluid", "thirion"])
def test_engine_multiscale_and_refine_matches_oracle(reg, params, niter):
    dimx, dimy = 192, 160
    R, T = S.make_pair(dimx, dimy, "lattice", shift=(2.5, -1.5), smooth=True)
    mf, tf = run(32, False, (dimx, dimy), R, T, reg, params, niter, nscales=2, nrefine=2)
    want = oracle(32).register(R, T, reg, params, niter, nscales=2, nrefine=2, verbose=1)
    assert tf["total_iterations"] == len(want["err"])
    assert maxdiff(mf, want["motion"]) <= 1e-3


def test_divide_by_zero_is_reported_like_the_reference():
    """coord2d::operator/ throws on a zero divisor (src/coord2d.h:95-100): flat, equal images hit it in Demons."""
    R = np.full((48, 64), 0.5)
    of.set_strict(False, 32)
    with of.Session((64, 48), [5], 0, of.THIRION, [1.0, 0.25, 1.5, 1.5, 5, 0], bits=32) as s:
        s.set_images(R, R)
        with pytest.raises(of.OF2DError) as e:
            s.estimate()
        assert e.value.code == 3 and "Divide by zero" in e.value.msg
