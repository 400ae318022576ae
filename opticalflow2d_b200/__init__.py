"""opticalflow2d_b200 -- B200-native (sm_100a) implementation of the per-iteration registration
solve of tjwdraper/OpticalFlow2d behind the reference's own API.

Layers (all native; Python only binds them for tests and benchmarks):

    lib/libof2d_cuda.so      hand-written CUDA kernels, C ABI in include/of2d_cuda.h
    lib/libof2d_host32.so    C++ classes with the reference's names (Image, Motion, Kernel,
    lib/libof2d_host64.so    ImageRegistration{OpticalFlow,Demons,Fluid}, solvers) + `mexFunction`,
                             C entry points in include/of2d_host.h (32 = float fields, 64 = fp64 mode)

Python surface:

    OpticalFlow2d(bits)      the Octave-facing function `OpticalFlow2d(...)` with its five call shapes
                             (WrapperOpticalFlow2d.cpp:18-155), driven through mexFunction
    Session(...)             the same classes without the MEX singleton (bench / batch use)
    cuda()                   raw access to the kernel ABI (kernel-level parity tests)

Importing the package never touches the GPU; the first call that needs a device raises if there is
none -- there is no CPU fallback anywhere in the product path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np

from . import _ffi

DIFFUSION, CURVATURE, ELASTIC, THIRION, DIFFEOMORPHIC, FLUID = range(6)
METHOD_NAMES = {0: "diffusion", 1: "curvature", 2: "elastic", 3: "thirion", 4: "diffeomorphic", 5: "fluid"}

__all__ = ["OpticalFlow2d", "Session", "Batch", "shard_pairs", "cuda", "host", "OF2DError", "DIFFUSION", "CURVATURE", "ELASTIC", "THIRION",
           "DIFFEOMORPHIC", "FLUID", "METHOD_NAMES"]


class OF2DError(RuntimeError):
    """A status code crossed the C ABI. code 2 = std::invalid_argument in the reference's terms,
    3 = std::runtime_error (divide by zero, mexErrMsgTxt, CUDA failure)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[status {code}] {msg}")
        self.code = code
        self.msg = msg


_cuda_lib = None
_host_libs: Dict[int, C.CDLL] = {}


def cuda() -> C.CDLL:
    """libof2d_cuda.so with every prototype of include/of2d_cuda.h typed."""
    global _cuda_lib
    if _cuda_lib is None:
        _cuda_lib = _ffi.load("libof2d_cuda.so", "of2d_cuda.h")
    return _cuda_lib


def host(bits: int = 32) -> C.CDLL:
    """libof2d_host{32,64}.so with every prototype of include/of2d_host.h typed."""
    if bits not in (32, 64):
        raise ValueError("bits must be 32 or 64")
    if bits not in _host_libs:
        cuda()   # fail with the clearer message if the kernel library itself is absent
        lib = _ffi.load(f"libof2d_host{bits}.so", "of2d_host.h")
        assert lib.of2d_host_real_bits() == bits
        _host_libs[bits] = lib
    return _host_libs[bits]


def loaded_libraries():
    out = []
    if _cuda_lib is not None:
        out.append(_cuda_lib._of2d_path)
    out += [l._of2d_path for l in _host_libs.values()]
    return out


def _check(lib, status: int):
    if status != 0:
        raise OF2DError(status, lib.of2d_host_last_error().decode(errors="replace"))


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _read_trace(lib, session) -> dict:
    nlev = lib.of2d_trace_num_levels(session)
    levels = []
    for k in range(nlev):
        sc, rf, it, nr = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        lib.of2d_trace_level_info(session, k, C.byref(sc), C.byref(rf), C.byref(it), C.byref(nr))
        series = []
        for which in range(5):
            n = lib.of2d_trace_level_series(session, k, which, None, 0)
            buf = np.zeros(max(n, 1))
            lib.of2d_trace_level_series(session, k, which, _ptr(buf), n)
            series.append(buf[:n].copy())
        levels.append({"scale": sc.value, "refine": rf.value, "iterations": it.value, "err": series[0],
                       "regrid_iter": series[1].astype(int), "regrid_minjac": series[2],
                       "fluid_maxabs": series[3], "fluid_dt": series[4]})
    return {"levels": levels, "total_iterations": int(lib.of2d_trace_total_iterations(session))}


class _Mx:
    """An in-process mxArray (real double, column-major), as Octave would hand to a MEX file."""

    def __init__(self, lib, values=None, dims: Optional[Sequence[int]] = None, handle=None):
        self.lib = lib
        if handle is not None:
            self.handle = handle
            return
        values = _f64(values)
        dims = list(dims) if dims is not None else [1, values.size]
        arr = (C.c_size_t * len(dims))(*dims)
        self.handle = lib.of2d_mx_create(len(dims), arr)
        C.memmove(lib.of2d_mx_data(self.handle), values.ctypes.data, values.nbytes)

    def numpy(self) -> np.ndarray:
        n = self.lib.of2d_mx_numel(self.handle)
        out = np.empty(n)
        C.memmove(out.ctypes.data, self.lib.of2d_mx_data(self.handle), out.nbytes)
        return out

    def free(self):
        if self.handle:
            self.lib.of2d_mx_free(self.handle)
            self.handle = None


class OpticalFlow2d:
    """The Octave function `OpticalFlow2d(...)` of the reference (one live registration object per
    process, WrapperOpticalFlow2d.cpp:13-16).  Images are numpy arrays of shape (dimy, dimx): C order
    with x fastest is byte-identical to MATLAB's column-major dimx x dimy array.

        of = OpticalFlow2d()
        of.init((dimx, dimy), niter, nscales, reg, regparams, nrefine, verbose)   # call shape (0, 8)
        of.register(Iref, Imov)                                                    # (0, 2)
        motion = of.motion()                # (dimy, dimx, 2)                      # (1, 0)
        Ireg = of.warp(Imov)                                                       # (1, 1)
        of.close()                                                                 # (0, 0)
    """

    def __init__(self, bits: int = 32):
        self.lib = host(bits)
        self.bits = bits
        self.shape = None

    def call(self, nlhs: int, args: Sequence[_Mx]):
        """Raw mexFunction(nlhs, plhs, nrhs, prhs)."""
        prhs = (C.c_void_p * max(len(args), 1))(*[a.handle for a in args])
        plhs = (C.c_void_p * 1)()
        _check(self.lib, self.lib.of2d_mex_call(nlhs, plhs, len(args), prhs))
        return _Mx(self.lib, handle=plhs[0]) if nlhs == 1 else None

    def init(self, dims, niter, nscales: int, reg: int, regparams, nrefine: int = 1, verbose: int = 0, nparams: Optional[int] = None):
        dimx, dimy = dims
        regparams = list(regparams)
        args = [_Mx(self.lib, [dimx, dimy]), _Mx(self.lib, list(niter)), _Mx(self.lib, [nscales]), _Mx(self.lib, [reg]),
                _Mx(self.lib, regparams if regparams else [0.0]), _Mx(self.lib, [len(regparams) if nparams is None else nparams]),
                _Mx(self.lib, [nrefine]), _Mx(self.lib, [verbose])]
        try:
            self.call(0, args)
        finally:
            for a in args:
                a.free()
        self.shape = (dimy, dimx)

    def register(self, Iref, Imov):
        Iref, Imov = _f64(Iref), _f64(Imov)
        dimy, dimx = Iref.shape
        args = [_Mx(self.lib, Iref, [dimx, dimy]), _Mx(self.lib, Imov, [dimx, dimy])]
        try:
            self.call(0, args)
        finally:
            for a in args:
                a.free()

    def motion(self) -> np.ndarray:
        out = self.call(1, [])
        try:
            planar = out.numpy().reshape(2, *self.shape)
        finally:
            out.free()
        return np.stack([planar[0], planar[1]], axis=-1)

    def warp(self, img) -> np.ndarray:
        img = _f64(img)
        dimy, dimx = img.shape
        arg = _Mx(self.lib, img, [dimx, dimy])
        try:
            out = self.call(1, [arg])
        finally:
            arg.free()
        try:
            return out.numpy().reshape(dimy, dimx)
        finally:
            out.free()

    def close(self):
        self.call(0, [])

    def trace(self) -> dict:
        return _read_trace(self.lib, None)


class Session:
    """ImageRegistration{OpticalFlow,Demons,Fluid} without the MEX singleton (src/ImageRegistration.h:14-29)."""

    def __init__(self, dims, niter, nscales: int, reg: int, regparams, nrefine: int = 1, verbose: int = 0, bits: int = 32):
        self.lib = host(bits)
        self.bits = bits
        dimx, dimy = dims
        self.shape = (dimy, dimx)
        niter_a = (C.c_int * (nscales + 1))(*[int(v) for v in niter])
        params = _f64(list(regparams) if len(regparams) else [0.0])
        self.handle = C.c_void_p()
        _check(self.lib, self.lib.of2d_session_create(dimx, dimy, nscales, niter_a, nrefine, reg, _ptr(params), len(regparams), verbose,
                                                      C.byref(self.handle)))

    def set_images(self, Iref, Imov):
        Iref, Imov = _f64(Iref), _f64(Imov)
        assert Iref.shape == self.shape and Imov.shape == self.shape
        _check(self.lib, self.lib.of2d_session_set_images(self.handle, _ptr(Iref), _ptr(Imov)))

    def set_images_raw(self, Iref_ptr: int, Imov_ptr: int):
        """Host pointers to dimx*dimy doubles each (e.g. pinned torch tensors)."""
        _check(self.lib, self.lib.of2d_session_set_images(self.handle, C.c_void_p(Iref_ptr), C.c_void_p(Imov_ptr)))

    def reset(self):
        """Extension: cold start (the reference warm-starts a second estimate from the previous result)."""
        _check(self.lib, self.lib.of2d_session_reset(self.handle))

    def estimate(self):
        _check(self.lib, self.lib.of2d_session_estimate(self.handle))

    def motion(self) -> np.ndarray:
        planar = np.zeros((2,) + self.shape)
        _check(self.lib, self.lib.of2d_session_get_motion(self.handle, _ptr(planar)))
        return np.stack([planar[0], planar[1]], axis=-1)

    def motion_raw(self, out_ptr: int):
        _check(self.lib, self.lib.of2d_session_get_motion(self.handle, C.c_void_p(out_ptr)))

    def motion_native(self) -> np.ndarray:
        out = np.zeros(self.shape + (2,), dtype=np.float32 if self.bits == 32 else np.float64)
        _check(self.lib, self.lib.of2d_session_get_motion_aos(self.handle, _ptr(out)))
        return out

    def warp(self, img) -> np.ndarray:
        img = _f64(img)
        out = np.zeros(self.shape)
        _check(self.lib, self.lib.of2d_session_warp(self.handle, _ptr(img), _ptr(out)))
        return out

    def trace(self) -> dict:
        return _read_trace(self.lib, self.handle)

    @staticmethod
    def register_many_raw(sessions, Iref_ptrs, Imov_ptrs, out_ptrs):
        """of2d_sessions_register: set_images + estimate + motion of several distinct sessions (same precision) in one call,
        the copies of the neighbouring jobs under each solve.  Host pointers as in set_images_raw / motion_raw."""
        n = len(sessions)
        lib = sessions[0].lib
        hs = (C.c_void_p * n)(*[s.handle for s in sessions])
        a = (C.c_void_p * n)(*[C.c_void_p(p) for p in Iref_ptrs])
        b = (C.c_void_p * n)(*[C.c_void_p(p) for p in Imov_ptrs])
        o = (C.c_void_p * n)(*[C.c_void_p(p) for p in out_ptrs])
        _check(lib, lib.of2d_sessions_register(hs, n, a, b, o))

    def close(self):
        if self.handle:
            self.lib.of2d_session_destroy(self.handle)
            self.handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class Batch:
    """`batch` independent pairs registered together (extension, include/of2d_host.h of2d_batch_*).
    Images: (batch, dimy, dimx) float64; motion: (batch, dimy, dimx, 2)."""

    def __init__(self, dims, batch: int, niter: int, reg: int, regparams, nrefine: int = 1, wave: int = 0, bits: int = 32, frames: int = 1, devices=None):
        """frames > 1: cine chains (frame-major batch, frame f warm-starts from frame f - 1: of2d_batch_create_chain);
        devices = [d0, d1, ...]: the pairs are sharded over these GPUs inside this process (of2d_batch_create_multi)."""
        self.lib = host(bits)
        dimx, dimy = dims
        self.shape = (batch, dimy, dimx)
        params = _f64(list(regparams) if len(regparams) else [0.0])
        self.handle = C.c_void_p()
        if frames > 1:
            assert devices is None, "cine chains run on one device"
            _check(self.lib, self.lib.of2d_batch_create_chain(dimx, dimy, batch, frames, niter, nrefine, reg, _ptr(params), len(regparams), C.byref(self.handle)))
        elif devices is not None:
            dev = np.ascontiguousarray(np.asarray(devices, dtype=np.int32))
            _check(self.lib, self.lib.of2d_batch_create_multi(dimx, dimy, batch, niter, nrefine, reg, _ptr(params), len(regparams), wave, _ptr(dev), len(dev), C.byref(self.handle)))
        else:
            _check(self.lib, self.lib.of2d_batch_create(dimx, dimy, batch, niter, nrefine, reg, _ptr(params), len(regparams), wave, C.byref(self.handle)))

    def register(self, Iref, Imov) -> np.ndarray:
        """Streamed protocol (of2d_batch_register): copies of the neighbouring waves run under the solve of the current one."""
        Iref, Imov = _f64(Iref), _f64(Imov)
        assert Iref.shape == self.shape and Imov.shape == self.shape
        b, dimy, dimx = self.shape
        planar = np.zeros((b, 2, dimy, dimx))
        _check(self.lib, self.lib.of2d_batch_register(self.handle, _ptr(Iref), _ptr(Imov), _ptr(planar)))
        return np.ascontiguousarray(np.moveaxis(planar, 1, -1))

    def register_raw(self, Iref_ptr: int, Imov_ptr: int, out_ptr: int):
        _check(self.lib, self.lib.of2d_batch_register(self.handle, C.c_void_p(Iref_ptr), C.c_void_p(Imov_ptr), C.c_void_p(out_ptr)))

    def shards(self):
        out = []
        for k in range(int(self.lib.of2d_batch_num_shards(self.handle))):
            d, lo, hi = C.c_int(), C.c_int(), C.c_int()
            _check(self.lib, self.lib.of2d_batch_shard_info(self.handle, k, C.byref(d), C.byref(lo), C.byref(hi)))
            out.append((d.value, lo.value, hi.value))
        return out

    def set_images(self, Iref, Imov):
        Iref, Imov = _f64(Iref), _f64(Imov)
        assert Iref.shape == self.shape and Imov.shape == self.shape
        _check(self.lib, self.lib.of2d_batch_set_images(self.handle, _ptr(Iref), _ptr(Imov)))

    def set_images_raw(self, Iref_ptr: int, Imov_ptr: int):
        _check(self.lib, self.lib.of2d_batch_set_images(self.handle, C.c_void_p(Iref_ptr), C.c_void_p(Imov_ptr)))

    def estimate(self):
        _check(self.lib, self.lib.of2d_batch_estimate(self.handle))

    def motion(self) -> np.ndarray:
        b, dimy, dimx = self.shape
        planar = np.zeros((b, 2, dimy, dimx))
        _check(self.lib, self.lib.of2d_batch_get_motion(self.handle, _ptr(planar)))
        return np.ascontiguousarray(np.moveaxis(planar, 1, -1))

    def motion_raw(self, out_ptr: int):
        _check(self.lib, self.lib.of2d_batch_get_motion(self.handle, C.c_void_p(out_ptr)))

    def iterations(self):
        b = self.shape[0]
        it = np.zeros(b, dtype=np.int32)
        rg = np.zeros(b, dtype=np.int32)
        _check(self.lib, self.lib.of2d_batch_iterations(self.handle, _ptr(it), _ptr(rg)))
        return it, rg

    def wave(self) -> int:
        return int(self.lib.of2d_batch_wave(self.handle))

    def close(self):
        if self.handle:
            self.lib.of2d_batch_destroy(self.handle)
            self.handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def shard_pairs(total: int, world: int, rank: int):
    """Contiguous shard [lo, hi) of `total` pairs for `rank` of `world` processes (one per GPU); the
    solve has no exchange step, so this partition is the whole multi-GPU protocol."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def set_strict(strict: bool, bits: int = 32):
    """strict=True: every kernel reproduces the reference's unfused arithmetic bit for bit."""
    lib = host(bits)
    _check(lib, lib.of2d_host_set_strict(int(strict)))


MATH_LEVELS = {"strict": 0, "exact": 1, "relaxed": 2}


def set_math(level, bits: int = 32):
    """Arithmetic level: 'strict' (0, the reference's loop literally), 'exact' (1, the engine with unfused arithmetic),
    'relaxed' (2, default: the engine with FMA contraction / approximate division / equivalent shortcuts)."""
    lib = host(bits)
    _check(lib, lib.of2d_host_set_math(int(MATH_LEVELS.get(level, level))))


def get_math(bits: int = 32) -> int:
    return int(host(bits).of2d_host_get_math())


def set_stream(cuda_stream: int, bits: int = 32):
    lib = host(bits)
    _check(lib, lib.of2d_host_set_stream(C.c_void_p(cuda_stream)))


def synchronize(bits: int = 32):
    lib = host(bits)
    _check(lib, lib.of2d_host_sync())


def profile_enable(on: bool, bits: int = 32):
    """Per-kernel CUDA-event timing of the engine kernels (bench.py's roofline leg)."""
    lib = host(bits)
    _check(lib, lib.of2d_host_profile_enable(int(on)))


def profile_read(bits: int = 32) -> dict:
    """{kernel: (launches, total_ms)} since profiling was enabled or last read."""
    lib = host(bits)
    buf = C.create_string_buffer(1 << 16)
    _check(lib, lib.of2d_host_profile_read(buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split()
        out[name] = (int(cnt), float(ms))
    return out


def launch_count(bits: int = 32) -> int:
    return int(host(bits).of2d_host_launch_count())
