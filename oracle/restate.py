"""ctypes binding of oracle/_ref/csc_restate.so (our CPU restatement, oracle/restate/).
TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "csc_restate.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "restate", "csc_restate.cpp")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            r = subprocess.run(["make", "-C", _HERE, "restate"], capture_output=True, text=True)
            if r.returncode:
                raise RuntimeError("restatement build failed:\n" + r.stdout + r.stderr)
        _lib = C.CDLL(_SO)
        _lib.rs_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_void_p,
                                   C.c_size_t, C.POINTER(C.c_size_t)]
        _lib.rs_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    return _lib


def _u8(a):
    if isinstance(a, (bytes, bytearray)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


def encode(data, dict_size, level, read_size=0, flt_mask=7, with_props=True):
    a = _u8(data)
    out = np.empty(a.size + a.size // 4 + (1 << 17), dtype=np.uint8)
    n = C.c_size_t(0)
    r = lib().rs_encode(a.ctypes.data, a.size, dict_size, level, read_size, flt_mask, int(with_props), out.ctypes.data,
                        out.size, C.byref(n))
    if r < 0:
        raise RuntimeError(f"rs_encode failed: {r}")
    return out[: n.value].copy()


def decode(stream, raw_size_hint):
    a = _u8(stream)
    out = np.empty(max(raw_size_hint, 16) + 64, dtype=np.uint8)
    n = C.c_size_t(0)
    r = lib().rs_decode(a.ctypes.data, a.size, out.ctypes.data, out.size, C.byref(n))
    if r < 0:
        raise RuntimeError(f"rs_decode failed: {r}")
    return out[: n.value].copy()
