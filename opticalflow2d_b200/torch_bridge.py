"""Kernel-level access for tests and benchmarks: torch owns device memory and the stream
("plumbing"), the kernels are called through the C ABI of include/of2d_cuda.h.

    dev = Device(strict=True)
    out = torch.empty_like(img)
    dev.call("warp2d", img.dtype, dimx, dimy, 1, img, u, out)     # -> of2d_warp2d_f32 / _f64
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import cuda


class KernelError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[status {code}] {msg}")
        self.code = code


class Device:
    def __init__(self, device: int | None = None, strict: bool = False):
        if not torch.cuda.is_available():
            raise RuntimeError("opticalflow2d_b200 needs a CUDA device (no CPU fallback)")
        self.lib = cuda()
        self.device = torch.cuda.current_device() if device is None else device
        self.ctx = C.c_void_p()
        self._check(self.lib.of2d_ctx_create(self.device, C.byref(self.ctx)))
        self.use_torch_stream()
        self.set_strict(strict)

    def _check(self, st):
        if st != 0:
            raise KernelError(st, self.lib.of2d_last_error().decode(errors="replace"))

    def use_torch_stream(self):
        self._check(self.lib.of2d_ctx_set_stream(self.ctx, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def set_strict(self, strict: bool):
        self._check(self.lib.of2d_ctx_set_fast_math(self.ctx, 0 if strict else 1))

    def sync(self):
        self._check(self.lib.of2d_ctx_sync(self.ctx))

    def launches(self) -> int:
        return int(self.lib.of2d_ctx_launch_count(self.ctx))

    @staticmethod
    def _arg(a):
        if isinstance(a, torch.Tensor):
            assert a.is_contiguous()
            return C.c_void_p(a.data_ptr())
        if isinstance(a, np.ndarray):
            assert a.flags["C_CONTIGUOUS"]
            return a.ctypes.data_as(C.c_void_p)
        if isinstance(a, np.floating):
            return float(a)
        if isinstance(a, np.integer):
            return int(a)
        return a

    def call(self, name: str, dtype, *args):
        suffix = "_f32" if dtype in (torch.float32, np.float32) else "_f64"
        fn = getattr(self.lib, "of2d_" + name + suffix)
        self._check(fn(self.ctx, *[self._arg(a) for a in args]))

    def call_plain(self, name: str, *args):
        fn = getattr(self.lib, "of2d_" + name)
        self._check(fn(self.ctx, *[self._arg(a) for a in args]))

    def close(self):
        if self.ctx:
            self.lib.of2d_ctx_destroy(self.ctx)
            self.ctx = C.c_void_p()


def to_dev(a: np.ndarray, dtype=None) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()
