"""ctypes bindings for the CHECKER libraries (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package opticalflow2d_b200 never does.

Two families with identical entry points (oracle/ref_shim.cpp and oracle/of2d_oracle.c):
  * RefLib("ref", 32|64)    -> oracle/_ref/libof2d_ref{32,64}.so : the reference's own
                               sources compiled unchanged (64 = float->double sed build)
  * RefLib("oracle", 32|64) -> oracle/liboracle_f{32,64}.so      : plain-C restatement

Array conventions: images are numpy arrays of shape (dimy, dimx) (C order, so x is the
fastest index, idx = i + j*dimx as in src/Field.tpp:13); motion fields are (dimy, dimx, 2)
array-of-structs {x, y}.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

DIFFUSION, CURVATURE, ELASTIC, THIRION, DIFFEOMORPHIC, FLUID = range(6)


class RefError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[status {code}] {msg}")
        self.code = code
        self.msg = msg


def lib_path(kind: str, bits: int) -> str:
    if kind == "ref":
        return os.path.join(HERE, "_ref", f"libof2d_ref{bits}.so")
    if kind == "oracle":
        return os.path.join(HERE, f"liboracle_f{bits}.so")
    raise ValueError(kind)


class RefLib:
    def __init__(self, kind: str = "ref", bits: int = 32):
        self.kind, self.bits = kind, bits
        self.path = lib_path(kind, bits)
        if not os.path.exists(self.path):
            raise FileNotFoundError(f"{self.path} missing: run `make -C oracle`")
        self.lib = C.CDLL(self.path, mode=os.RTLD_LOCAL if hasattr(os, "RTLD_LOCAL") else 0)
        self.real = np.float32 if bits == 32 else np.float64
        self.creal = C.c_float if bits == 32 else C.c_double
        self.prefix = "of2d_ref_" if kind == "ref" else "of2d_oracle_"
        assert self._f("sizeof_real")() == bits // 8
        self._f("last_error").restype = C.c_char_p

    # -- helpers -----------------------------------------------------------------------------
    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def _check(self, status: int):
        if status != 0:
            raise RefError(status, self._f("last_error")().decode())

    def _r(self, a) -> np.ndarray:
        return np.ascontiguousarray(a, dtype=self.real)

    @staticmethod
    def _p(a: np.ndarray):
        return a.ctypes.data_as(C.c_void_p)

    # -- trace (captured mexPrintf varargs) ----------------------------------------------------
    def trace_reset(self):
        self.lib.of2d_trace_reset()

    def trace(self, which: int):
        n = self.lib.of2d_trace_count(which)
        a = np.zeros(n)
        b = np.zeros(n)
        if n:
            self.lib.of2d_trace_get(which, self._p(a), self._p(b), n)
        return a, b

    # -- full protocol -------------------------------------------------------------------------
    def register(self, Iref, Imov, reg: int, regparams: Sequence[float], niter: Sequence[int],
                 nscales: int = 0, nrefine: int = 1, verbose: int = 1):
        """Replays test_opticalflow2d.m:42-59 through mexFunction. Returns dict with motion
        (dimy, dimx, 2) float64, warped image, and the captured traces."""
        Iref = np.ascontiguousarray(Iref, dtype=np.float64)
        Imov = np.ascontiguousarray(Imov, dtype=np.float64)
        dimy, dimx = Iref.shape
        n = dimx * dimy
        niter_d = np.asarray(niter, dtype=np.float64)
        assert niter_d.size == nscales + 1
        params = np.asarray(regparams, dtype=np.float64)
        motion = np.zeros(2 * n)
        warped = np.zeros(n)
        self.trace_reset()
        st = self._f("mex_register")(C.c_int(dimx), C.c_int(dimy), C.c_int(nscales), self._p(niter_d),
                                     C.c_int(nrefine), C.c_int(reg), self._p(params), C.c_int(params.size),
                                     C.c_int(verbose), self._p(Iref), self._p(Imov), self._p(motion),
                                     self._p(warped))
        self._check(st)
        it, err = self.trace(0)
        rg_it, rg_mj = self.trace(1)
        fl_ma, fl_dt = self.trace(2)
        planar = motion.reshape(2, dimy, dimx)
        return {
            "motion": np.stack([planar[0], planar[1]], axis=-1),
            "warped": warped.reshape(dimy, dimx),
            "err_iter": it.astype(int), "err": err,
            "regrid_iter": rg_it.astype(int), "regrid_minjac": rg_mj,
            "fluid_maxabs": fl_ma, "fluid_dt": fl_dt,
        }

    def mex_badcall(self, nlhs: int, nrhs: int) -> int:
        return self._f("mex_badcall")(C.c_int(nlhs), C.c_int(nrhs))

    # -- primitives ----------------------------------------------------------------------------
    def set_image(self, img64):
        img64 = np.ascontiguousarray(img64, dtype=np.float64)
        dimy, dimx = img64.shape
        out = np.zeros((dimy, dimx), dtype=self.real)
        self._check(self._f("set_image")(dimx, dimy, self._p(img64), self._p(out)))
        return out

    def copy_motion_to_input(self, u):
        u = self._r(u)
        dimy, dimx, _ = u.shape
        out = np.zeros((2, dimy, dimx), dtype=np.float64)
        self._check(self._f("copy_motion_to_input")(dimx, dimy, self._p(u), self._p(out)))
        return out

    def warp2d(self, img, u):
        img = self._r(img).copy()
        u = self._r(u)
        dimy, dimx = img.shape
        self._check(self._f("warp2d")(dimx, dimy, self._p(img), self._p(u)))
        return img

    def accumulate(self, u, v):
        u = self._r(u).copy()
        v = self._r(v)
        dimy, dimx, _ = u.shape
        self._check(self._f("accumulate")(dimx, dimy, self._p(u), self._p(v)))
        return u

    def gaussian_kernel(self, w: int, sigma: float):
        out = np.zeros((w, w), dtype=np.float64)
        self._check(self._f("gaussian_kernel")(C.c_int(w), self.creal(sigma), self._p(out)))
        return out

    def convolute_motion(self, u, w: int, sigma: float):
        u = self._r(u).copy()
        dimy, dimx, _ = u.shape
        self._check(self._f("convolute_motion")(dimx, dimy, self._p(u), C.c_int(w), self.creal(sigma)))
        return u

    # -- the rest of the public Image / Motion / Kernel surface --
    def image_stats(self, img):
        img = self._r(img)
        dimy, dimx = img.shape
        s, hi, lo = self.creal(0), self.creal(0), self.creal(0)
        self._check(self._f("image_stats")(dimx, dimy, self._p(img), C.byref(s), C.byref(hi), C.byref(lo)))
        return s.value, hi.value, lo.value

    def image_normalize(self, img):
        img = self._r(img).copy()
        dimy, dimx = img.shape
        self._check(self._f("image_normalize")(dimx, dimy, self._p(img)))
        return img

    def boundary_conditions(self, u, kind: int):
        u = self._r(u).copy()
        dimy, dimx, _ = u.shape
        self._check(self._f("boundary_conditions")(dimx, dimy, C.c_int(kind), self._p(u)))
        return u

    def average_kernel(self, w: int):
        out = np.zeros((w, w), dtype=np.float64)
        self._check(self._f("average_kernel")(C.c_int(w), self._p(out)))
        return out

    def convolute_image(self, img, w: int, sigma: float):
        """sigma > 0: Gaussian kernel; sigma <= 0: Kernel::set_average"""
        img = self._r(img).copy()
        dimy, dimx = img.shape
        self._check(self._f("convolute_image")(dimx, dimy, self._p(img), C.c_int(w), self.creal(sigma)))
        return img

    def exp(self, u):
        u = self._r(u).copy()
        dimy, dimx, _ = u.shape
        self._check(self._f("exp")(dimx, dimy, self._p(u)))
        return u

    def norm_maxabs(self, u):
        u = self._r(u)
        dimy, dimx, _ = u.shape
        a, b = self.creal(0), self.creal(0)
        self._check(self._f("norm_maxabs")(dimx, dimy, self._p(u), C.byref(a), C.byref(b)))
        return a.value, b.value

    def jacobian(self, u):
        u = self._r(u)
        dimy, dimx, _ = u.shape
        jac = np.zeros((dimy, dimx), dtype=self.real)
        mj = self.creal(0)
        self._check(self._f("jacobian")(dimx, dimy, self._p(u), self._p(jac), C.byref(mj)))
        return jac, mj.value

    def derivatives(self, Iref, Imov):
        Iref, Imov = self._r(Iref), self._r(Imov)
        dimy, dimx = Iref.shape
        g = np.zeros((dimy, dimx, 2), dtype=self.real)
        it = np.zeros((dimy, dimx), dtype=self.real)
        self._check(self._f("derivatives")(dimx, dimy, self._p(Iref), self._p(Imov), self._p(g), self._p(it)))
        return g, it

    def image_resample(self, img, out_shape, up: bool, out_init=None):
        img = self._r(img)
        iny, inx = img.shape
        outy, outx = out_shape
        out = np.zeros((outy, outx), dtype=self.real) if out_init is None else self._r(out_init).copy()
        self._check(self._f("image_resample")(inx, iny, self._p(img), outx, outy, self._p(out), int(up)))
        return out

    def motion_resample(self, u, out_shape, up: bool, out_init=None):
        u = self._r(u)
        iny, inx, _ = u.shape
        outy, outx = out_shape
        out = np.zeros((outy, outx, 2), dtype=self.real) if out_init is None else self._r(out_init).copy()
        self._check(self._f("motion_resample")(inx, iny, self._p(u), outx, outy, self._p(out), int(up)))
        return out

    def logger(self, useq):
        useq = self._r(useq)
        nseq, dimy, dimx, _ = useq.shape
        err = np.zeros(nseq, dtype=self.real)
        self._check(self._f("logger")(dimx, dimy, self._p(useq), nseq, self._p(err)))
        return err

    def solver_steps(self, reg: int, params: Sequence[float], Iref, Imov, u, nsteps: int):
        Iref, Imov = self._r(Iref), self._r(Imov)
        u = self._r(u).copy()
        dimy, dimx = Iref.shape
        p = np.asarray(params, dtype=self.real)
        self._check(self._f("solver_steps")(C.c_int(reg), self._p(p), C.c_int(p.size), dimx, dimy,
                                            self._p(Iref), self._p(Imov), self._p(u), C.c_int(nsteps)))
        return u

    def dct1d(self, x, kind: int):
        x = np.ascontiguousarray(x, dtype=np.float64).copy()
        self.lib.of2d_standin_dct1d(self._p(x), C.c_int(x.size), C.c_int(kind))
        return x


_cache = {}


def get(kind: str = "ref", bits: int = 32) -> RefLib:
    key = (kind, bits)
    if key not in _cache:
        _cache[key] = RefLib(kind, bits)
    return _cache[key]


def available(kind: str = "ref", bits: int = 32) -> bool:
    return os.path.exists(lib_path(kind, bits))
