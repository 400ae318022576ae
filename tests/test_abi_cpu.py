"""CPU suite: the native libraries build for sm_100a, load, and export every symbol the public
headers declare; the product path fails loudly without a GPU and never touches the oracle."""
import ctypes as C
import os
import re
import subprocess

import pytest

import opticalflow2d_b200 as of
from opticalflow2d_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_libraries_build_and_export_declared_symbols(built_libs):
    for lib, header in (("libof2d_cuda.so", "of2d_cuda.h"), ("libof2d_host32.so", "of2d_host.h"), ("libof2d_host64.so", "of2d_host.h")):
        path = os.path.join(_ffi.LIBDIR, lib)
        assert os.path.exists(path)
        exported = set(re.findall(r" T (\w+)", subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout))
        declared = _ffi.declared_symbols(header)
        assert len(declared) > 20
        missing = [s for s in declared if s not in exported]
        assert not missing, f"{lib} does not export {missing}"
    # the MEX entry point itself (WrapperOpticalFlow2d.cpp:18-20), unmangled as Octave expects
    for bits in (32, 64):
        out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(_ffi.LIBDIR, f"libof2d_host{bits}.so")], capture_output=True, text=True).stdout
        assert re.search(r" T mexFunction\b", out)


def test_kernels_are_compiled_for_sm_100a(built_libs):
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(_ffi.LIBDIR, "libof2d_cuda.so")], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_ctypes_prototypes_cover_every_entry_point(built_libs):
    lib = of.cuda()
    for name in _ffi.declared_symbols("of2d_cuda.h"):
        assert getattr(lib, name).argtypes is not None
    assert of.host(32).of2d_host_real_bits() == 32
    assert of.host(64).of2d_host_real_bits() == 64


def test_f32_and_f64_entry_points_come_in_pairs():
    names = set(_ffi.declared_symbols("of2d_cuda.h"))
    f32 = {n[:-4] for n in names if n.endswith("_f32")}
    f64 = {n[:-4] for n in names if n.endswith("_f64")}
    assert f32 - f64 == set() and f64 - f32 <= {"of2d_dct2d"}


def test_product_sources_never_reference_the_oracle():
    """The oracle is test infrastructure: nothing under opticalflow2d_b200/ or include/ may import,
    include, link or dlopen it."""
    bad = []
    for base in ("opticalflow2d_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if os.path.basename(dirpath) in ("build", "lib", "__pycache__"):
                continue
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".tpp")):
                    continue
                text = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"(from|import)\s+oracle\b|liboracle|of2d_oracle|oracle/_ref|libof2d_ref", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_no_cpu_fallback_without_a_gpu(built_libs):
    ctx = C.c_void_p()
    lib = of.cuda()
    assert lib.of2d_ctx_create(0, C.byref(ctx)) == 1          # OF2D_ERR_CUDA
    assert b"cuda" in lib.of2d_last_error().lower()
    with pytest.raises(of.OF2DError) as e:
        of.Session((32, 32), [5], 0, 0, [0.5])
    assert e.value.code == 3 and "no CPU path" in e.value.msg
    o = of.OpticalFlow2d()
    with pytest.raises(of.OF2DError):
        o.init((32, 32), [5], 0, 0, [0.5])


def test_missing_library_raises(monkeypatch, tmp_path):
    monkeypatch.setattr(_ffi, "LIBDIR", str(tmp_path))
    with pytest.raises(_ffi.NativeLibraryMissing):
        _ffi.load("libof2d_cuda.so", "of2d_cuda.h")
