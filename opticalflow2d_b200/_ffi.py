"""ctypes plumbing: loads the in-tree native libraries and types every entry point from the C
prototypes in include/*.h (so the headers stay the single source of truth for the C ABI).

There is no fallback: a missing library raises, and every compute entry point fails when no CUDA
device is present.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict, List, Tuple

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIBDIR = os.path.join(PKG, "lib")
INCLUDE = os.path.join(ROOT, "include")

_SCALARS = {
    "int": C.c_int, "unsigned": C.c_uint, "unsigned int": C.c_uint, "long": C.c_long, "size_t": C.c_size_t,
    "float": C.c_float, "double": C.c_double, "uint64_t": C.c_uint64, "unsigned long long": C.c_ulonglong,
    "void": None,
}

_PROTO = re.compile(r"^\s*([A-Za-z_][\w\s\*]*?)\s*\b(of2d_\w+)\s*\(([^;{]*)\)\s*;", re.M)


def _ctype(decl: str):
    decl = decl.strip()
    if "*" in decl:
        if re.match(r"^(const\s+)?char\s*\*$", decl):
            return C.c_char_p
        return C.c_void_p
    decl = re.sub(r"\bconst\b", "", decl).strip()
    if decl in _SCALARS:
        return _SCALARS[decl]
    raise ValueError(f"unmapped C type {decl!r}")


def parse_header(path: str) -> Dict[str, Tuple[object, List[object]]]:
    """name -> (restype, argtypes) for every `of2d_*` prototype of a header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    out = {}
    for ret, name, args in _PROTO.findall(text):
        ret = ret.replace("extern", "").strip()
        if ret.startswith("typedef"):
            continue
        argtypes = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                # drop the parameter name (last identifier) unless the declaration is a bare type
                m = re.match(r"^(.*?[\*\s])(\w+)$", a)
                typ = m.group(1) if m else a
                argtypes.append(_ctype(typ))
        out[name] = (_ctype(ret), argtypes)
    return out


def declared_symbols(header: str) -> List[str]:
    return sorted(parse_header(os.path.join(INCLUDE, header)).keys())


class NativeLibraryMissing(RuntimeError):
    pass


def load(libname: str, header: str) -> C.CDLL:
    path = os.path.join(LIBDIR, libname)
    if not os.path.exists(path):
        raise NativeLibraryMissing(
            f"{path} is missing: build it with `python -m opticalflow2d_b200.build` "
            "(the CUDA extension is mandatory; there is no CPU fallback)")
    lib = C.CDLL(path, mode=getattr(os, "RTLD_LOCAL", 0))
    for name, (restype, argtypes) in parse_header(os.path.join(INCLUDE, header)).items():
        fn = getattr(lib, name)   # AttributeError here means the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    lib._of2d_path = path
    return lib
