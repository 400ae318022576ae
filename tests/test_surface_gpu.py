"""The rest of the reference's public Image / Motion / Kernel surface (methods no driver calls) on the device,
against the oracle restatement, which tests/test_oracle_cpu.py pins to the COMPILED reference bit for bit (-m gpu):

  Image::sum / max / min      src/Image.cpp:78-104      of2d_image_stats_*        (sum: parallel order, 1e-6 relative)
  Image::normalize            src/Image.cpp:107-116     of2d_image_normalize_*
  Motion::Neumann_/Dirichlet_ src/Motion.cpp:181-251    of2d_boundary_conditions_*  (including the y-extent-as-x-index corner)
  Image::convolute            src/Image.cpp:184-187     of2d_convolute_image_*    (reference: uninitialised accumulator, UB; here 0)
  Kernel::set_average         src/Kernel.cpp:75-82      host class

both through the C ABI of libof2d_cuda (strict context) and through the host classes (of2d_host_image_op / _motion_boundary /
_kernel, which build Image / Motion / Kernel objects and call the methods)."""
import ctypes as C

import numpy as np
import pytest
import torch

import opticalflow2d_b200 as of
from gpu_common import NP, TD, device, maxdiff, oracle
from opticalflow2d_b200 import synthetic as S
from opticalflow2d_b200.torch_bridge import to_dev

pytestmark = pytest.mark.gpu
SIZES = [(64, 48), (97, 33), (33, 97), (256, 256)]
BITS = [32, 64]


def _img(dimx, dimy, bits, seed=3):
    rng = np.random.default_rng(seed)
    return (S.make_pair(dimx, dimy, "lattice")[0] + 0.2 * rng.uniform(-1, 1, (dimy, dimx))).astype(NP[bits])


@pytest.mark.parametrize("bits", BITS)
@pytest.mark.parametrize("dimx,dimy", SIZES)
def test_image_stats_and_normalize(bits, dimx, dimy):
    dev, orc = device(), oracle(bits)
    img = _img(dimx, dimy, bits)
    s, hi, lo = orc.image_stats(img)
    d = to_dev(img)
    cs, chi, clo = (C.c_float if bits == 32 else C.c_double)(), (C.c_float if bits == 32 else C.c_double)(), (C.c_float if bits == 32 else C.c_double)()
    dev.call("image_stats", TD[bits], dimx * dimy, d, C.byref(cs), C.byref(chi), C.byref(clo))
    assert chi.value == hi and clo.value == lo
    assert abs(cs.value - s) <= (2e-5 if bits == 32 else 1e-12) * abs(s)      # the reference's sequential accumulation vs a tree
    want = orc.image_normalize(img)
    dev.call("image_normalize", TD[bits], dimx * dimy, NP[bits](lo), NP[bits](hi), d)
    assert np.array_equal(d.cpu().numpy(), want)


@pytest.mark.parametrize("bits", BITS)
@pytest.mark.parametrize("dimx,dimy", SIZES + [(5, 9), (9, 5)])
@pytest.mark.parametrize("kind", [0, 1], ids=["neumann", "dirichlet"])
def test_boundary_conditions(bits, dimx, dimy, kind):
    dev, orc = device(), oracle(bits)
    u = S.random_motion(dimx, dimy, 2.0, 5, False).astype(NP[bits])
    want = orc.boundary_conditions(u, kind)
    d = to_dev(u)
    dev.call("boundary_conditions", TD[bits], dimx, dimy, kind, d)
    assert np.array_equal(d.cpu().numpy(), want)


@pytest.mark.parametrize("bits", BITS)
@pytest.mark.parametrize("dimx,dimy", SIZES[:3] + [(7, 9)])
@pytest.mark.parametrize("w,sigma", [(5, 1.5), (3, 0.8), (3, -1.0), (5, -1.0), (4, 1.0)], ids=["g5", "g3", "avg3", "avg5", "g4"])
def test_convolute_image(bits, dimx, dimy, w, sigma):
    dev, orc = device(), oracle(bits)
    img = _img(dimx, dimy, bits, 4)
    want = orc.convolute_image(img, w, sigma)
    k = orc.gaussian_kernel(w, sigma) if sigma > 0 else orc.average_kernel(w)
    d_in = to_dev(img)
    d_out = torch.empty_like(d_in)
    dev.call("convolute_image", TD[bits], dimx, dimy, 1, d_in, d_out, np.ascontiguousarray(k), w, w)
    # the reference promotes every tap product to double (float field x double weight); the kernel multiplies by the weight rounded to `real`: a few ulp
    assert maxdiff(d_out.cpu().numpy(), want) <= (5e-7 if bits == 32 else 9e-16) * max(1.0, float(np.abs(want).max()))


@pytest.mark.parametrize("bits", BITS)
def test_host_classes_image_motion_kernel_surface(bits):
    """Image::sum/max/min/normalize/convolute, Motion::*_boundaryconditions, Kernel::set_gaussian/set_average through the host classes."""
    lib, orc = of.host(bits), oracle(bits)
    dimx, dimy = 97, 65
    img = _img(dimx, dimy, bits).astype(np.float64)     # values exactly representable in `real`
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    sc = np.zeros(3)
    out = np.zeros_like(img)
    assert lib.of2d_host_image_op(0, dimx, dimy, p(img), None, p(sc), 0, C.c_double(0.0)) == 0, lib.of2d_host_last_error()
    s, hi, lo = orc.image_stats(img)
    assert sc[1] == hi and sc[2] == lo and abs(sc[0] - s) <= (2e-5 if bits == 32 else 1e-12) * abs(s)
    assert lib.of2d_host_image_op(1, dimx, dimy, p(img), p(out), None, 0, C.c_double(0.0)) == 0
    assert np.array_equal(out.astype(NP[bits]), orc.image_normalize(img))
    for w, sigma in ((5, 1.5), (3, -1.0)):
        assert lib.of2d_host_image_op(2, dimx, dimy, p(img), p(out), None, w, C.c_double(sigma)) == 0
        want = orc.convolute_image(img, w, sigma)
        assert maxdiff(out, want) <= (5e-7 if bits == 32 else 9e-16) * max(1.0, float(np.abs(want).max()))
    u = S.random_motion(dimx, dimy, 2.0, 7, False).astype(NP[bits]).astype(np.float64)
    uo = np.zeros_like(u)
    for kind in (0, 1):
        assert lib.of2d_host_motion_boundary(kind, dimx, dimy, p(u), p(uo)) == 0
        assert np.array_equal(uo.astype(NP[bits]), orc.boundary_conditions(u, kind))
    for w in (3, 5, 7):
        k = np.zeros((w, w))
        assert lib.of2d_host_kernel(1, w, C.c_double(0.0), p(k)) == 0
        assert np.array_equal(k, orc.average_kernel(w))
        assert lib.of2d_host_kernel(0, w, C.c_double(1.5), p(k)) == 0
        assert np.array_equal(k, orc.gaussian_kernel(w, 1.5))
