"""ctypes binding of lib/libcloudatlas_b200.so (declared in include/cloudatlas_b200.h).

The library is the product; this module only loads it and checks status codes.  There is no
Python or CPU fallback: if the shared library is missing, importing the package fails.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CLOUDATLAS_B200_LIB: another build of the same library (tools/build_variant.sh: tuning experiments)
LIB_PATH = os.environ.get("CLOUDATLAS_B200_LIB") or os.path.join(_HERE, "lib", "libcloudatlas_b200.so")


class CloudAtlasError(RuntimeError):
    """Raised where the Julia wrapper would call error(msg)."""

    def __init__(self, code, msg):
        super().__init__(f"[cloudatlas_b200 status {code}] {msg}")
        self.code = code


if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU fallback)")

lib = C.CDLL(LIB_PATH)

c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
c_i32p = C.POINTER(C.c_int32)
vp = C.c_void_p

lib.ca_last_error.restype = C.c_char_p
lib.ca_launch_count.restype = C.c_int64


def check(status: int):
    if status != 0:
        raise CloudAtlasError(status, lib.ca_last_error().decode("utf-8", "replace"))


def _sig(name, *argtypes):
    fn = getattr(lib, name)
    fn.argtypes = list(argtypes)
    fn.restype = C.c_int32
    return fn


def exported_symbols():
    """Names declared in include/cloudatlas_b200.h (parsed), for the ABI test."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "cloudatlas_b200.h")
    text = open(hdr).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ca_[a-z0-9_]+)\s*\(", text)))
