"""Per-iteration parity (-m gpu): one or more get_update() steps of each solver through the kernel
ABI against the oracle's solver_steps (same state in, compare state out).  Strict mode, bit-exact,
except Curvature where the device DCT and the CPU stand-in are different double-precision FFTs:
tolerance 2e-6 px per step in fp32 (one float ulp at |u| < 8) and 1e-12 in fp64."""
import ctypes as C

import numpy as np
import pytest
import torch

from gpu_common import NP, TD, device, maxdiff, oracle, pair
from opticalflow2d_b200 import synthetic as S
from opticalflow2d_b200.torch_bridge import to_dev

pytestmark = pytest.mark.gpu
BITS = [32, 64]
SIZES = [(64, 48), (97, 35), (200, 130)]


def _setup(bits, dimx, dimy, kind="lattice"):
    orc = oracle(bits)
    R, T = pair(dimx, dimy, kind, smooth=True, sigma_b=6.0)
    R, T = R.astype(NP[bits]), T.astype(NP[bits])
    g, it = orc.derivatives(R, T)
    u0 = S.random_motion(dimx, dimy, 0.3, 21, True).astype(NP[bits])
    return orc, R, T, g, it, u0


@pytest.mark.parametrize("bits", BITS)
@pytest.mark.parametrize("dimx,dimy", SIZES)
def test_diffusion_steps(bits, dimx, dimy):
    dev = device()
    orc, R, T, g, it, u0 = _setup(bits, dimx, dimy)
    want = orc.solver_steps(0, [0.5], R, T, u0, 5)
    a, b = to_dev(u0), torch.empty((dimy, dimx, 2), dtype=TD[bits], device="cuda")
    dg, dit = to_dev(g), to_dev(it)
    for _ in range(5):
        dev.call("diffusion_step", TD[bits], dimx, dimy, 1, a, b, dg, dit, NP[bits](0.5), None)
        a, b = b, a
    assert np.array_equal(a.cpu().numpy(), want)


@pytest.mark.parametrize("bits", BITS)
def test_diffusion_divide_by_zero_is_reported(bits):
    """alpha = 0 on a flat image: the reference throws 'Divide by zero exception' (coord2d.h:95-100)."""
    dev = device()
    orc = oracle(bits)
    dimx, dimy = 32, 32
    flat = np.ones((dimy, dimx), dtype=NP[bits])
    from oracle.refapi import RefError
    with pytest.raises(RefError) as e:
        orc.solver_steps(0, [0.0], flat, flat, np.zeros((dimy, dimx, 2), dtype=NP[bits]), 1)
    assert e.value.code == 3
    z2 = torch.zeros((dimy, dimx, 2), dtype=TD[bits], device="cuda")
    z1 = torch.zeros((dimy, dimx), dtype=TD[bits], device="cuda")
    status = (C.c_uint * 1)(0)
    dev.call("diffusion_step", TD[bits], dimx, dimy, 1, z2, torch.empty_like(z2), z2.clone(), z1, NP[bits](0.0), status)
    assert status[0] & 1


@pytest.mark.parametrize("bits", BITS)
@pytest.mark.parametrize("dimx,dimy", SIZES + [(3, 3), (34, 5), (33, 70), (130, 40)])
@pytest.mark.parametrize("params", [(1.0, 0.25, 0.66), (0.5, 0.0, 1.2)])
def test_elastic_sor_sweep_is_exact_lexicographic(bits, dimx, dimy, params):
    dev = device()
    orc, R, T, g, it, u0 = _setup(bits, dimx, dimy)
    mu, la, om = [NP[bits](p) for p in params]
    want = orc.solver_steps(2, list(params), R, T, u0, 3)
    d_u = to_dev(u0)
    dg, dit = to_dev(g), to_dev(it)
    for _ in range(3):
        dev.call("elastic_step", TD[bits], dimx, dimy, 1, d_u, dg, dit, mu, la, om)
    assert np.array_equal(d_u.cpu().numpy(), want)


@pytest.mark.parametrize("bits", BITS)
def test_elastic_batched_matches_single(bits):
    dev = device()
    dimx, dimy, batch = 70, 50, 5
    orc = oracle(bits)
    us, gs, its, wants = [], [], [], []
    for k in range(batch):
        R, T = S.batch_pair(k, dimx, dimy)
        R, T = R.astype(NP[bits]), T.astype(NP[bits])
        g, it = orc.derivatives(R, T)
        u0 = S.random_motion(dimx, dimy, 0.2, 30 + k, True).astype(NP[bits])
        us.append(u0); gs.append(g); its.append(it)
        wants.append(orc.solver_steps(2, [1.0, 0.25], R, T, u0, 2))
    d_u, dg, dit = to_dev(np.stack(us)), to_dev(np.stack(gs)), to_dev(np.stack(its))
    for _ in range(2):
        dev.call("elastic_step", TD[bits], dimx, dimy, batch, d_u, dg, dit, NP[bits](1.0), NP[bits](0.25), NP[bits](np.float32(0.66)))
    assert np.array_equal(d_u.cpu().numpy(), np.stack(wants))


@pytest.mark.parametrize("bits", BITS)
@pytest.mark.parametrize("dimx,dimy", SIZES)
def test_fluid_steps(bits, dimx, dimy):
    dev = device()
    orc, R, T, g, it, _ = _setup(bits, dimx, dimy)
    u0 = np.zeros((dimy, dimx, 2), dtype=NP[bits])
    orc.trace_reset()
    want = orc.solver_steps(5, [0.1, 0.0], R, T, u0, 6)
    ma_want, dt_want = orc.trace(2)
    d_u = to_dev(u0)
    d_v, d_r = torch.zeros_like(d_u), torch.zeros_like(d_u)
    dg, dit = to_dev(g), to_dev(it)
    hm, hd = np.zeros(1, dtype=NP[bits]), np.zeros(1, dtype=NP[bits])
    omega = NP[bits](0.66)   # OpticalFlowFluid.h:10: the default is the double literal 0.66 narrowed to the field type
    for k in range(6):
        dev.call("fluid_step", TD[bits], dimx, dimy, d_u, d_v, d_r, dg, dit, NP[bits](0.1), NP[bits](0.0), omega, hm, hd)
        assert float(hm[0]) == ma_want[k] and float(hd[0]) == dt_want[k]
    assert np.array_equal(d_u.cpu().numpy(), want)


@pytest.mark.parametrize("bits", BITS)
@pytest.mark.parametrize("dimx,dimy", SIZES)
def test_demons_force_fused_warp_gradient(bits, dimx, dimy):
    dev = device()
    orc, R, T, _, _, _ = _setup(bits, dimx, dimy)
    u = S.random_motion(dimx, dimy, 2.0, 22, True).astype(NP[bits])
    Iwar = orc.warp2d(T, u)
    g, it = orc.derivatives(R, Iwar)
    si, sx = NP[bits](1.0), NP[bits](0.25)
    den = g[..., 0] * g[..., 0] + g[..., 1] * g[..., 1] + it * it * (si * si) / (sx * sx)
    want = np.stack([g[..., 0] * it / den * NP[bits](-1), g[..., 1] * it / den * NP[bits](-1)], axis=-1)
    d_c = torch.empty((dimy, dimx, 2), dtype=TD[bits], device="cuda")
    dev.call("demons_force", TD[bits], dimx, dimy, 1, to_dev(R), to_dev(T), to_dev(u), d_c, si, sx, None)
    assert np.array_equal(d_c.cpu().numpy(), want)


@pytest.mark.parametrize("bits", BITS)
@pytest.mark.parametrize("reg,params", [(3, [1, 0.25, 1.5, 1.5, 5, 0]), (3, [1, 0.25, 1.5, 1.5, 5, 1]), (4, [1, 2.0, 1.5, 1.5, 5])])
def test_demons_get_update_chain(bits, reg, params):
    """force -> K_fluid -> (exp) -> compose/add -> K_diffusion, three iterations, as DemonsThirions.cpp:18-42."""
    dev = device()
    dimx, dimy = 96, 64
    orc, R, T, _, _, _ = _setup(bits, dimx, dimy)
    u0 = np.zeros((dimy, dimx, 2), dtype=NP[bits])
    want = orc.solver_steps(reg, params, R, T, u0, 3)
    kd = orc.gaussian_kernel(int(params[4]), params[2])
    kf = orc.gaussian_kernel(int(params[4]), params[3])
    w = int(params[4])
    dR, dT = to_dev(R), to_dev(T)
    u = to_dev(u0)
    c, s = torch.empty_like(u), torch.empty_like(u)
    for _ in range(3):
        dev.call("demons_force", TD[bits], dimx, dimy, 1, dR, dT, u, s, NP[bits](params[0]), NP[bits](params[1]), None)
        dev.call("convolute_motion", TD[bits], dimx, dimy, 1, s, c, kf, w, w)
        if reg == 4:
            dev.call("motion_exp", TD[bits], dimx, dimy, c, s, None)
        if reg == 4 or params[5] == 0:
            dev.call("compose", TD[bits], dimx, dimy, 1, u, c, s)
            u, s = s, u
        else:
            dev.call("axpy", TD[bits], dimx * dimy * 2, NP[bits](1), c, u)
        dev.call("convolute_motion", TD[bits], dimx, dimy, 1, u, s, kd, w, w)
        u, s = s, u
    assert np.array_equal(u.cpu().numpy(), want)


# non-powers of two: Bluestein (any n up to 4096, incl. primes and the demo's padded 278 = 2 x 139); 5000: direct sum
@pytest.mark.parametrize("n0,n1", [(8, 8), (64, 32), (256, 128), (12, 20), (2048, 4), (278, 256), (139, 97), (3, 5), (1000, 6), (7, 3000), (4, 4095), (3, 5000)])
@pytest.mark.parametrize("kind", [2, 3])
def test_dct2d_matches_fftw_definition(n0, n1, kind):
    """The device transform against scipy's DCT-II/III (norm=None), i.e. FFTW's REDFT10/REDFT01."""
    import scipy.fft
    dev = device()
    rng = np.random.default_rng(5)
    x = rng.standard_normal((n0, n1))
    want = scipy.fft.dctn(x, type=kind, norm=None)
    d = to_dev(x)
    dev.call_plain("dct2d_f64", n0, n1, kind, d)
    got = d.cpu().numpy()
    assert maxdiff(got, want) <= 1e-12 * np.abs(want).max()


@pytest.mark.parametrize("bits", BITS)
@pytest.mark.parametrize("dimx,dimy", [(64, 32), (128, 128), (48, 40), (64, 64), (256, 128), (64, 512), (1024, 64), (2048, 256), (128, 4096),
                                       (512, 512), (1024, 512), (512, 2048), (4096, 512), (2048, 1024),   # >= 512 both ways: register path (dct_reg.cuh)
                                       (278, 256), (256, 278), (139, 97), (600, 360)])                      # non-powers of two: Bluestein (test_opticalflow2d.m:14-20 pads to 278 x 256)
def test_curvature_steps(bits, dimx, dimy):
    dev = device()
    orc, R, T, g, it, _ = _setup(bits, dimx, dimy)
    u0 = np.zeros((dimy, dimx, 2), dtype=NP[bits])
    nsteps = 4
    want = orc.solver_steps(1, [0.25, 1.0], R, T, u0, nsteps)
    plan = C.c_void_p()
    dev.call_plain("curvature_plan_create", dimx, dimy, C.c_double(0.25), C.c_double(1.0), int(bits == 64), C.byref(plan))
    a, b = to_dev(u0), torch.empty((dimy, dimx, 2), dtype=TD[bits], device="cuda")
    dg, dit = to_dev(g), to_dev(it)
    fn = getattr(dev.lib, "of2d_curvature_step_f32" if bits == 32 else "of2d_curvature_step_f64")
    for _ in range(nsteps):
        st = fn(plan, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(dg.data_ptr()), C.c_void_p(dit.data_ptr()))
        assert st == 0, dev.lib.of2d_last_error()
        a, b = b, a
    got = a.cpu().numpy()
    dev.lib.of2d_curvature_plan_destroy(plan)
    tol = 2e-6 if bits == 32 else 1e-12
    assert maxdiff(got, want) <= tol
