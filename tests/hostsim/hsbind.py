"""ctypes binding to tests/hostsim/_build/hostsim.so -- the engine sources compiled for the
host (single-lane warp).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.abspath(os.path.join(_HERE, "..", ".."))
_SO = os.path.join(_HERE, "_build", "hostsim.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, "hostsim.cpp")] + [
        os.path.join(_ROOT, "csc_b200", "csrc", f) for f in os.listdir(os.path.join(_ROOT, "csc_b200", "csrc"))
        if f.endswith((".h", ".cuh"))]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
        return
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    cmd = ["g++", "-O2", "-g", "-Wall", "-Wno-unused", "-Wno-unknown-pragmas", "-fPIC", "-shared",
           "-I", os.path.join(_ROOT, "include"), "-o", _SO, os.path.join(_HERE, "hostsim.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("hostsim build failed:\n" + r.stderr)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.hs_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_uint32, C.c_int, C.c_int,
                                   C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t,
                                   C.POINTER(C.c_size_t)]
        _lib.hs_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        _lib.hs_analyze_blocks.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        _lib.hs_filter.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32]
        _lib.hs_props_init.argtypes = [C.c_uint32, C.c_int, C.c_void_p]
    return _lib


def _u8(a):
    if isinstance(a, (bytes, bytearray)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


def encode(data, dict_size, level, read_size=0, flt_mask=7, with_props=True, trace=False, cands=False):
    a = _u8(data)
    trace = trace or cands   # the engine always records candidate lists when a trace buffer is given
    out = np.empty(a.size * 2 + (1 << 18), dtype=np.uint8)
    n = C.c_size_t(0)
    tn = C.c_size_t(0)
    tr = np.zeros(((a.size + 1024) * (12 if cands else 1), 3), dtype=np.uint32) if trace else None
    r = lib().hs_encode(a.ctypes.data, a.size, dict_size, level, read_size, flt_mask, int(with_props),
                        out.ctypes.data, out.size, C.byref(n), tr.ctypes.data if trace else None,
                        tr.shape[0] if trace else 0, C.byref(tn))
    if r < 0:
        raise RuntimeError(f"hs_encode failed: {r}")
    res = out[: n.value].copy()
    return (res, tr[: tn.value]) if trace else res


def decode(stream, raw_size_hint):
    a = _u8(stream)
    out = np.empty(max(raw_size_hint, 16) + 64, dtype=np.uint8)
    n = C.c_size_t(0)
    r = lib().hs_decode(a.ctypes.data, a.size, out.ctypes.data, out.size - 64, C.byref(n))
    if r < 0:
        raise RuntimeError(f"hs_decode failed: {r}")
    return out[: n.value].copy()


def analyze_blocks(data):
    a = _u8(data)
    nb = (a.size + 8191) // 8192
    t = np.zeros(nb, dtype=np.uint32)
    b = np.zeros(nb, dtype=np.uint32)
    lib().hs_analyze_blocks(a.ctypes.data, a.size, t.ctypes.data, b.ctypes.data)
    return t, b


def filt(which, inverse, data, chn=0):
    a = _u8(data).copy()
    r = lib().hs_filter(which, int(inverse), a.ctypes.data, a.size, chn)
    return a, r


def props(dict_size, level):
    o = np.zeros(15, dtype=np.uint64)
    lib().hs_props_init(dict_size, level, o.ctypes.data)
    k = ["dict_size", "csc_blocksize", "raw_blocksize", "hash_bits", "hash_width", "bt_hash_bits", "bt_size",
         "bt_cyc", "good_len", "lz_mode", "DLTFilter", "TXTFilter", "EXEFilter", "est_mem", "sizeof"]
    return {kk: int(v) for kk, v in zip(k, o)}


class Stages:
    """Intermediate arrays of one single-stream encode (analyzer, filters, match-finder rows)."""

    def __init__(self, n, K):
        nblk = (n + 8191) // 8192 + 4
        self.binfo = np.zeros(nblk, np.uint32); self.bbpb = np.zeros(nblk, np.uint32)
        self.lz = np.zeros(n + 64, np.uint8); self.P6 = np.zeros(n * K + 8, np.uint32); self.L6 = np.zeros(n * K + 8, np.uint16)
        self.Z2 = np.zeros((n + 31) // 32 + 2, np.uint32); self.Z6 = np.zeros((n + 31) // 32 + 2, np.uint32)
        self.dict_ok = np.zeros(nblk, np.uint8); self.tail_dlt = np.zeros(64 * 5, np.uint32)

    def ptrs(self):
        return [a.ctypes.data for a in (self.binfo, self.bbpb, self.lz, self.P6, self.L6, self.Z2, self.Z6, self.dict_ok,
                                        self.tail_dlt)]


def encode_stages(data, dict_size, level):
    """hostsim encode + its intermediate arrays."""
    a = _u8(data)
    K = 16 if level == 5 else props(dict_size, level)["hash_width"]   # -m5: bounded-search rows (csc_defs.h kBtRowK)
    st = Stages(a.size, max(K, 1))
    l = lib()
    l.hs_debug_dump.argtypes = [C.c_void_p] * 9
    l.hs_debug_dump(*st.ptrs())
    try:
        out = encode(a, dict_size, level)
    finally:
        l.hs_debug_dump(*([None] * 9))
    return out, st
