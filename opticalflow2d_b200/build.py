"""Build recipe for the native libraries (in-tree, so the .so files travel with gpurun snapshots).

  libof2d_cuda.so     csrc/*.cu    hand-written sm_100a kernels behind include/of2d_cuda.h
  libof2d_host32.so   host/**.cpp  C++ mirror of the reference API (float fields) + MEX shim
  libof2d_host64.so   same sources with -DOF2D_REAL=double (fp64 mode)

`python -m opticalflow2d_b200.build [--force] [--verbose]`
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
LIBDIR = os.path.join(PKG, "lib")
OBJDIR = os.path.join(PKG, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fno-fast-math",
    "--expt-relaxed-constexpr",
    "-I" + os.path.join(ROOT, "include"),
]
# arithmetic: every translation unit reproduces the reference's unfused mul/add rounding and IEEE division ...
EXACT_FLAGS = ["-fmad=false"]
# ... except *_relaxed.cu (the second build of the iteration engine, arithmetic level 2)
RELAXED_FLAGS = ["-fmad=true", "-prec-div=false", "-prec-sqrt=false"]
CXX_FLAGS = ["-std=c++17", "-O2", "-fPIC", "-pthread", "-Wall", "-Wno-unused-function", "-I" + os.path.join(ROOT, "include"), "-I" + HOST, "-I" + os.path.join(HOST, "mex")]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: the CUDA path cannot be built (there is no CPU fallback)")
    return cand


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build step failed: " + " ".join(cmd[:3]) + " ...")
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr)


def build_cuda(force=False, verbose=False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(ROOT, "include", "*.h")))
    out = os.path.join(LIBDIR, "libof2d_cuda.so")
    objs, jobs = [], []
    for s in srcs:
        o = os.path.join(OBJDIR, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _newer(o, [s] + hdrs + ([s.replace("_relaxed.cu", ".cu")] if s.endswith("_relaxed.cu") else [])):
            jobs.append([nvcc] + NVCC_FLAGS + (RELAXED_FLAGS if s.endswith("_relaxed.cu") else EXACT_FLAGS) + os.environ.get("OF2D_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(lambda c: _run(c, verbose), jobs))
    if jobs or force or _newer(out, objs):
        _run([nvcc, "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"], verbose)
    return out


def build_host(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    srcs = sorted(glob.glob(os.path.join(HOST, "**", "*.cpp"), recursive=True))
    hdrs = sorted(glob.glob(os.path.join(HOST, "**", "*.h"), recursive=True)) + \
        sorted(glob.glob(os.path.join(HOST, "**", "*.tpp"), recursive=True)) + \
        sorted(glob.glob(os.path.join(ROOT, "include", "*.h")))
    outs = []
    cuda_lib = os.path.join(LIBDIR, "libof2d_cuda.so")
    for bits, real in ((32, "float"), (64, "double")):
        out = os.path.join(LIBDIR, f"libof2d_host{bits}.so")
        outs.append(out)
        if not srcs:
            continue
        if force or _newer(out, srcs + hdrs + [cuda_lib]):
            _run([cxx] + CXX_FLAGS + [f"-DOF2D_REAL={real}", f"-DOF2D_REAL_BITS={bits}"] + srcs +
                 ["-shared", "-Wl,-Bsymbolic", "-o", out, "-L" + LIBDIR, "-lof2d_cuda", "-Wl,-rpath,$ORIGIN"], verbose)
    return outs


def build_all(force=False, verbose=False):
    libs = [build_cuda(force, verbose)]
    libs += build_host(force, verbose)
    return libs


if __name__ == "__main__":
    print("\n".join(build_all(force="--force" in sys.argv, verbose="--verbose" in sys.argv)))
