"""Kernel-level parity (-m gpu): every primitive of include/of2d_cuda.h against the oracle's
restatement of the reference function it replaces.  In strict mode the kernels follow the
reference's operation order without FMA contraction, so the bar is BIT-EXACT for everything except
the float-accumulated reductions (Motion::norm / Logger), where the reference's sequential float sum
cannot be reproduced by a parallel reduction: tolerance 2e-4 relative."""
import ctypes as C

import numpy as np
import pytest
import torch

from gpu_common import NP, TD, device, maxdiff, oracle, pair
from opticalflow2d_b200 import synthetic as S
from opticalflow2d_b200.torch_bridge import to_dev

pytestmark = pytest.mark.gpu

SIZES = [(64, 48), (97, 33), (256, 256)]
BITS = [32, 64]


def _motion(dimx, dimy, amp, seed, bits, smooth=True):
    return S.random_motion(dimx, dimy, amp, seed, smooth).astype(NP[bits])


@pytest.mark.parametrize("bits", BITS)
@pytest.mark.parametrize("dimx,dimy", SIZES)
@pytest.mark.parametrize("amp", [0.7, 6.0, 80.0])
def test_warp2d(bits, dimx, dimy, amp):
    dev, orc = device(), oracle(bits)
    img = pair(dimx, dimy)[0].astype(NP[bits])
    u = _motion(dimx, dimy, amp, 1, bits)
    want = orc.warp2d(img, u)
    d_img, d_u = to_dev(img), to_dev(u)
    d_out = torch.empty_like(d_img)
    dev.call("warp2d", TD[bits], dimx, dimy, 1, d_img, d_u, d_out)
    assert np.array_equal(d_out.cpu().numpy(), want)


@pytest.mark.parametrize("bits", BITS)
@pytest.mark.parametrize("dimx,dimy", SIZES)
@pytest.mark.parametrize("amp", [0.4, 5.0, 70.0])
def test_compose(bits, dimx, dimy, amp):
    dev, orc = device(), oracle(bits)
    u = _motion(dimx, dimy, 2.0, 2, bits)
    v = _motion(dimx, dimy, amp, 3, bits)
    want = orc.accumulate(u, v)
    d_u, d_v = to_dev(u), to_dev(v)
    d_out = torch.empty_like(d_u)
    dev.call("compose", TD[bits], dimx, dimy, 1, d_u, d_v, d_out)
    assert np.array_equal(d_out.cpu().numpy(), want)


@pytest.mark.parametrize("bits", BITS)
@pytest.mark.parametrize("dimx,dimy", SIZES + [(7, 9)])
@pytest.mark.parametrize("w,sigma", [(5, 1.5), (3, 0.8), (7, 2.0), (4, 1.0), (9, 3.0)])
def test_convolute_motion_wraps_like_reference(bits, dimx, dimy, w, sigma):
    """Includes the linear-index wrap of Field.tpp:245-248, even widths and kernels wider than the image."""
    dev, orc = device(), oracle(bits)
    u = _motion(dimx, dimy, 3.0, 4, bits, smooth=False)
    want = orc.convolute_motion(u, w, sigma)
    k = orc.gaussian_kernel(w, sigma)
    d_u = to_dev(u)
    d_out = torch.empty_like(d_u)
    dev.call("convolute_motion", TD[bits], dimx, dimy, 1, d_u, d_out, k, w, w)
    assert np.array_equal(d_out.cpu().numpy(), want)


@pytest.mark.parametrize("bits", BITS)
def test_convolute_fast_mode_within_rounding(bits):
    dev, orc = device(strict=False), oracle(bits)
    dimx, dimy = 128, 96
    u = _motion(dimx, dimy, 3.0, 5, bits, smooth=False)
    want = orc.convolute_motion(u, 5, 1.5)
    k = orc.gaussian_kernel(5, 1.5)
    d_u = to_dev(u)
    d_out = torch.empty_like(d_u)
    dev.call("convolute_motion", TD[bits], dimx, dimy, 1, d_u, d_out, k, 5, 5)
    tol = 2e-6 if bits == 32 else 1e-14
    assert maxdiff(d_out.cpu().numpy(), want) <= tol


@pytest.mark.parametrize("bits", BITS)
@pytest.mark.parametrize("dimx,dimy", SIZES + [(2, 2)])
def test_derivatives(bits, dimx, dimy):
    dev, orc = device(), oracle(bits)
    R, T = pair(dimx, dimy)
    R, T = R.astype(NP[bits]), T.astype(NP[bits])
    g_want, it_want = orc.derivatives(R, T)
    d_g = torch.empty((dimy, dimx, 2), dtype=TD[bits], device="cuda")
    d_it = torch.empty((dimy, dimx), dtype=TD[bits], device="cuda")
    dev.call("derivatives", TD[bits], dimx, dimy, 1, to_dev(R), to_dev(T), d_g, d_it)
    assert np.array_equal(d_g.cpu().numpy(), g_want)
    assert np.array_equal(d_it.cpu().numpy(), it_want)


@pytest.mark.parametrize("bits", BITS)
@pytest.mark.parametrize("dimx,dimy", SIZES)
def test_jacobian_and_min(bits, dimx, dimy):
    dev, orc = device(), oracle(bits)
    u = _motion(dimx, dimy, 4.0, 6, bits)
    jac_want, min_want = orc.jacobian(u)
    d_jac = torch.empty((dimy, dimx), dtype=TD[bits], device="cuda")
    h_min = np.zeros(1, dtype=NP[bits])
    dev.call("jacobian", TD[bits], dimx, dimy, to_dev(u), d_jac, h_min)
    assert np.array_equal(d_jac.cpu().numpy(), jac_want)
    assert h_min[0] == NP[bits](min_want)


@pytest.mark.parametrize("bits", BITS)
@pytest.mark.parametrize("dimx,dimy", SIZES)
def test_norm_maxabs_logger(bits, dimx, dimy):
    dev, orc = device(), oracle(bits)
    n = dimx * dimy
    u = _motion(dimx, dimy, 2.5, 7, bits)
    norm_want, maxabs_want = orc.norm_maxabs(u)
    h = np.zeros(1, dtype=NP[bits])
    dev.call("motion_norm", TD[bits], n, to_dev(u), h)
    assert abs(h[0] - norm_want) <= 2e-4 * abs(norm_want)
    dev.call("motion_maxabs", TD[bits], n, to_dev(u), h)
    assert h[0] == NP[bits](maxabs_want)            # a max is order independent: exact, including the y-twice quirk
    assert abs(maxabs_want - np.sqrt(2.0) * np.abs(u[..., 1]).max()) < 1e-5 * maxabs_want

    seq = np.stack([_motion(dimx, dimy, 1.0 + 0.3 * k, 10 + k, bits) for k in range(4)])
    err_want = orc.logger(seq)
    d_prev = torch.zeros((dimy, dimx, 2), dtype=TD[bits], device="cuda")
    for k in range(4):
        dev.call("logger_update", TD[bits], n, to_dev(seq[k]), d_prev, h)
        assert abs(h[0] - err_want[k]) <= 3e-4 * abs(err_want[k]) + 1e-12
        assert np.array_equal(d_prev.cpu().numpy(), seq[k])
    assert err_want[0] == 0.0


@pytest.mark.parametrize("bits", BITS)
@pytest.mark.parametrize("amp", [0.05, 0.3, 1.7, 9.0])
def test_motion_exp_scaling_and_squaring(bits, amp):
    dev, orc = device(), oracle(bits)
    dimx, dimy = 96, 80
    u = _motion(dimx, dimy, amp, 8, bits)
    want = orc.exp(u)
    d_u = to_dev(u)
    d_tmp = torch.empty_like(d_u)
    ns = C.c_int(-1)
    dev.call("motion_exp", TD[bits], dimx, dimy, d_u, d_tmp, C.byref(ns))
    assert np.array_equal(d_u.cpu().numpy(), want)
    _, ma = orc.norm_maxabs(u)
    expect_ns = max(0, int(np.ceil(1 + np.log2(NP[bits](ma))))) if ma > 0 else 0
    assert ns.value == expect_ns
    if amp >= 1.7:
        assert ns.value > 0      # squaring actually exercised


@pytest.mark.parametrize("bits", BITS)
def test_motion_exp_of_zero_field(bits):
    dev = device()
    d_u = torch.zeros((16, 16, 2), dtype=TD[bits], device="cuda")
    ns = C.c_int(-1)
    dev.call("motion_exp", TD[bits], 16, 16, d_u, torch.empty_like(d_u), C.byref(ns))
    assert ns.value == 0 and float(d_u.abs().max()) == 0.0


@pytest.mark.parametrize("bits", BITS)
@pytest.mark.parametrize("shape_in,shape_out", [((64, 48), (32, 24)), ((100, 72), (25, 18)), ((65, 33), (32, 16)), ((50, 50), (50, 50))])
def test_resample_image_and_motion(bits, shape_in, shape_out):
    dev, orc = device(), oracle(bits)
    (ix, iy), (ox, oy) = shape_in, shape_out
    img = pair(ix, iy)[0].astype(NP[bits])
    u = _motion(ix, iy, 3.0, 9, bits)
    # down
    want = orc.image_resample(img, (oy, ox), up=False)
    d_out = torch.zeros((oy, ox), dtype=TD[bits], device="cuda")
    dev.call("downsample", TD[bits], 1, ix, iy, to_dev(img), ox, oy, d_out)
    assert np.array_equal(d_out.cpu().numpy(), want)
    want_m = orc.motion_resample(u, (oy, ox), up=False)
    d_m = torch.zeros((oy, ox, 2), dtype=TD[bits], device="cuda")
    dev.call("downsample", TD[bits], 2, ix, iy, to_dev(u), ox, oy, d_m)
    dev.call("scale_xy", TD[bits], ox * oy, NP[bits](ox) / NP[bits](ix), NP[bits](oy) / NP[bits](iy), d_m)
    assert np.array_equal(d_m.cpu().numpy(), want_m)
    # up (from the small grid back to the large one)
    small, small_m = want, want_m
    want_up = orc.image_resample(small, (iy, ix), up=True)
    d_up = torch.zeros((iy, ix), dtype=TD[bits], device="cuda")
    dev.call("upsample", TD[bits], 1, ox, oy, to_dev(small), ix, iy, d_up)
    assert np.array_equal(d_up.cpu().numpy(), want_up)
    want_upm = orc.motion_resample(small_m, (iy, ix), up=True)
    d_upm = torch.zeros((iy, ix, 2), dtype=TD[bits], device="cuda")
    dev.call("upsample", TD[bits], 2, ox, oy, to_dev(small_m), ix, iy, d_upm)
    dev.call("scale_xy", TD[bits], ix * iy, NP[bits](ix) / NP[bits](ox), NP[bits](iy) / NP[bits](oy), d_upm)
    assert np.array_equal(d_upm.cpu().numpy(), want_upm)


@pytest.mark.parametrize("bits", BITS)
def test_io_casts(bits):
    dev, orc = device(), oracle(bits)
    dimx, dimy = 70, 41
    img64 = pair(dimx, dimy)[0] * np.pi
    want = orc.set_image(img64)
    d_out = torch.empty((dimy, dimx), dtype=TD[bits], device="cuda")
    dev.call("image_from_double", TD[bits], dimx * dimy, to_dev(img64), d_out)
    assert np.array_equal(d_out.cpu().numpy(), want)
    u = _motion(dimx, dimy, 2.0, 11, bits)
    want_p = orc.copy_motion_to_input(u)
    d_p = torch.empty((2, dimy, dimx), dtype=torch.float64, device="cuda")
    dev.call("motion_to_planar_double", TD[bits], dimx * dimy, to_dev(u), d_p)
    assert np.array_equal(d_p.cpu().numpy(), want_p)


def test_out_of_place_contract_is_enforced():
    dev = device()
    t = torch.zeros((8, 8), dtype=torch.float32, device="cuda")
    u = torch.zeros((8, 8, 2), dtype=torch.float32, device="cuda")
    from opticalflow2d_b200.torch_bridge import KernelError
    with pytest.raises(KernelError) as e:
        dev.call("warp2d", torch.float32, 8, 8, 1, t, u, t)
    assert e.value.code == 2
