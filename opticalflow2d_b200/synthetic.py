"""Closed-form synthetic image pairs.

The reference ships no data (img/ is git-ignored, test_opticalflow2d.m:8-9), so every
configuration in BASELINE.json is exercised on these generated pairs.  Images are float64
numpy arrays of shape (dimy, dimx): x (index i) is the fastest axis, matching the reference's
column-major MATLAB layout idx = i + j*dimx (src/Field.tpp:13).

    tex(i,j)     = 0.05*(sin(0.11 i) + cos(0.07 j))        -- no exactly flat region, so the
                                                              divide-by-zero throw of coord2d.h:95
                                                              is never hit by benchmark inputs
    blob(i,j)    = exp(-((i-nx/2)^2+(j-ny/2)^2)/(2 (n/8)^2))
    lattice(i,j) = exp(-(fx^2+fy^2)/(2 sb^2)), fx = mod(i,64)-32, fy = mod(j,64)-32
    R = base + tex,  T(i,j) = R(i-sx, j-sy)  (evaluated in closed form, no resampling)
"""
from __future__ import annotations

import numpy as np

_MASK = (1 << 64) - 1


def splitmix64(state: int):
    """One step of splitmix64; returns (new_state, output)."""
    state = (state + 0x9E3779B97F4A7C15) & _MASK
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
    return state, z ^ (z >> 31)


def uniforms(seed: int, n: int) -> np.ndarray:
    out = np.empty(n)
    s = seed & _MASK
    for k in range(n):
        s, z = splitmix64(s)
        out[k] = (z >> 11) * (1.0 / (1 << 53))
    return out


def _grid(dimx: int, dimy: int):
    i = np.arange(dimx, dtype=np.float64)[None, :]
    j = np.arange(dimy, dtype=np.float64)[:, None]
    return i, j


def _tex(i, j):
    return 0.05 * (np.sin(0.11 * i) + np.cos(0.07 * j))


def _blob(i, j, dimx, dimy):
    n = min(dimx, dimy)
    return np.exp(-((i - dimx / 2.0) ** 2 + (j - dimy / 2.0) ** 2) / (2.0 * (n / 8.0) ** 2))


def _lattice(i, j, sb):
    fx = np.mod(i, 64.0) - 32.0
    fy = np.mod(j, 64.0) - 32.0
    return np.exp(-(fx * fx + fy * fy) / (2.0 * sb * sb))


def make_pair(dimx: int, dimy: int, kind: str = "blob", shift=(1.5, -0.75), smooth: bool = False,
              sigma_b: float = 8.0, noise_seed: int | None = None):
    """Returns (Iref, Imov) float64 (dimy, dimx). kind in {"blob", "lattice"}."""
    i, j = _grid(dimx, dimy)

    def base(ii, jj):
        b = _blob(ii, jj, dimx, dimy) if kind == "blob" else _lattice(ii, jj, sigma_b)
        return b + _tex(ii, jj)

    sx, sy = shift
    if smooth:
        sxf = sx * (1.0 + 0.5 * np.sin(2.0 * np.pi * j / dimy))
        syf = sy * (1.0 + 0.5 * np.cos(2.0 * np.pi * i / dimx))
    else:
        sxf, syf = sx, sy
    R = base(i + 0.0 * j, j + 0.0 * i)
    T = base(i - sxf + 0.0 * j, j - syf + 0.0 * i)
    if noise_seed is not None:
        u = uniforms(noise_seed, 2 * dimx * dimy)
        R = R + 1e-3 * (2.0 * u[: dimx * dimy].reshape(dimy, dimx) - 1.0)
        T = T + 1e-3 * (2.0 * u[dimx * dimy:].reshape(dimy, dimx) - 1.0)
    return np.ascontiguousarray(R), np.ascontiguousarray(T)


def batch_pair(k: int, dimx: int = 512, dimy: int = 512):
    """Pair k of the batched configuration (C5): per-pair shift and lattice width."""
    u = uniforms(0xB200 + k, 3)
    return make_pair(dimx, dimy, kind="lattice", shift=(-2.0 + 4.0 * u[0], -2.0 + 4.0 * u[1]),
                     smooth=True, sigma_b=5.0 + 3.0 * u[2])


def random_motion(dimx: int, dimy: int, amp: float, seed: int, smooth: bool = True) -> np.ndarray:
    """A (dimy, dimx, 2) float64 displacement field for primitive-level tests."""
    rng = np.random.default_rng(seed)
    if not smooth:
        return amp * (2.0 * rng.random((dimy, dimx, 2)) - 1.0)
    i, j = _grid(dimx, dimy)
    ph = rng.random(4) * 2.0 * np.pi
    ux = amp * np.sin(2.0 * np.pi * i / dimx * 1.5 + ph[0]) * np.cos(2.0 * np.pi * j / dimy + ph[1])
    uy = amp * np.cos(2.0 * np.pi * i / dimx + ph[2]) * np.sin(2.0 * np.pi * j / dimy * 2.0 + ph[3])
    out = np.stack([ux + 0.0 * j, uy + 0.0 * i], axis=-1)
    out += 0.05 * amp * (2.0 * rng.random((dimy, dimx, 2)) - 1.0)
    return np.ascontiguousarray(out)
