"""Shared helpers of the -m gpu parity tests: the CUDA path is compared with the CPU oracle
(oracle/of2d_oracle.c, pinned against the compiled reference) on identical seeded inputs."""
import numpy as np
import torch

import opticalflow2d_b200 as of
from opticalflow2d_b200 import synthetic as S
from opticalflow2d_b200.torch_bridge import Device, to_dev
from oracle import refapi

NP = {32: np.float32, 64: np.float64}
TD = {32: torch.float32, 64: torch.float64}

_devices = {}


def device(strict=True) -> Device:
    if strict not in _devices:
        _devices[strict] = Device(strict=strict)
    d = _devices[strict]
    d.use_torch_stream()
    return d


def oracle(bits):
    return refapi.get("oracle", bits)


def pair(dimx, dimy, kind="lattice", **kw):
    return S.make_pair(dimx, dimy, kind, **kw)


def maxdiff(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))
