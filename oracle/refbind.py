"""ctypes bindings to oracle/_ref/libcsc_ref.so (the UNMODIFIED reference, built by
oracle/Makefile).  TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py's CPU
baseline / reference arm and __graft_entry__.smoke() -- never by csc_b200."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
REF_SRC = "/root/reference/src"


def build(verbose: bool = False) -> None:
    """Build oracle/_ref (needs /root/reference; on the GPU box the prebuilt .so travels)."""
    if not os.path.isdir(REF_SRC):
        return
    r = subprocess.run(["make", "-C", _HERE, "ref"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)
    # the reference's own command-line tool, once on the reference library and once relinked against the
    # product (needs csc_b200/csrc/libcsc_b200.so; skipped quietly when the product is not built yet)
    if os.path.exists(os.path.join(_HERE, "..", "csc_b200", "csrc", "libcsc_b200.so")):
        r = subprocess.run(["make", "-C", _HERE, "cli"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle/_ref cli build failed:\n" + r.stdout + r.stderr)


def cli(which: str) -> str:
    """Path of oracle/_ref/csc_cli_ref or csc_cli_b200 (libcsc/csc.cpp, unmodified), built on demand."""
    path = os.path.join(_REF_DIR, which if which.startswith("csarc") else "csc_cli_" + which)
    if not os.path.exists(path):
        build()
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    return path


def _load(name: str):
    path = os.path.join(_REF_DIR, name)
    if not os.path.exists(path):
        build()
    if not os.path.exists(path):
        raise FileNotFoundError(path + " (oracle/_ref not built and /root/reference absent)")
    return C.CDLL(path)


_lib = None
_tlib = None


def available() -> bool:
    return os.path.exists(os.path.join(_REF_DIR, "libcsc_ref.so")) or os.path.isdir(REF_SRC)


def lib():
    global _lib
    if _lib is None:
        _lib = _load("libcsc_ref.so")
        _sig(_lib)
    return _lib


def tlib():
    global _tlib
    if _tlib is None:
        _tlib = _load("libcsc_ref_trace.so")
        _sig(_tlib)
        _tlib.ref_trace_begin.argtypes = [C.c_void_p, C.c_size_t]
        _tlib.ref_trace_begin_cands.argtypes = [C.c_void_p, C.c_size_t]
        _tlib.ref_trace_end.restype = C.c_size_t
    return _tlib


def _sig(l):
    l.ref_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_uint32, C.c_int, C.c_int,
                             C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    l.ref_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    l.ref_props_init.argtypes = [C.c_uint32, C.c_int, C.c_void_p]
    l.ref_analyze_blocks.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    l.ref_dlt_bpb.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    l.ref_dlt_bpb.restype = C.c_uint32
    l.ref_filter.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32]
    l.ref_run_chunks.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, C.c_int]


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


def out_bound(n: int) -> int:
    return int(n + n // 4 + (1 << 17))


def encode(data, dict_size: int, level: int, read_size: int = 0, flt_mask: int = 7,
           with_props: bool = True, trace: bool = False, cands: bool = False):
    """Reference stream for one handle; dict_size is the argument of CSCEncProps_Init.
    trace: also return the packet trace; cands: the trace additionally carries the candidate list of every
    MatchFinder::FindMatch / FindMatchWithPrice call (oracle/ref_mftrace.cpp)."""
    trace = trace or cands
    a = _u8(data)
    l = tlib() if trace else lib()
    out = np.empty(out_bound(a.size), dtype=np.uint8)
    n = C.c_size_t(0)
    tr = None
    if trace:
        tr = np.zeros((max(16, a.size + 1024) * (12 if cands else 1), 3), dtype=np.uint32)
        (l.ref_trace_begin_cands if cands else l.ref_trace_begin)(tr.ctypes.data, tr.shape[0])
    r = l.ref_encode(a.ctypes.data, a.size, dict_size, level, read_size, flt_mask, int(with_props),
                     out.ctypes.data, out.size, C.byref(n))
    if trace:
        cnt = l.ref_trace_end()
        tr = tr[:cnt]
    if r < 0:
        raise RuntimeError(f"ref_encode failed: {r}")
    res = out[: n.value].copy()
    return (res, tr) if trace else res


def encode_parts(data, dict_size: int, level: int, part_sizes) -> np.ndarray:
    """Reference stream when the input is fed through several CSCEnc_Encode calls on ONE handle."""
    a = _u8(data)
    assert sum(part_sizes) == a.size
    l = lib()
    l.ref_encode_parts.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                   C.POINTER(C.c_size_t)]
    ps = (C.c_size_t * len(part_sizes))(*part_sizes)
    out = np.empty(out_bound(a.size), dtype=np.uint8)
    n = C.c_size_t(0)
    r = l.ref_encode_parts(a.ctypes.data, ps, len(part_sizes), dict_size, level, 1, out.ctypes.data, out.size, C.byref(n))
    if r < 0:
        raise RuntimeError(f"ref_encode_parts failed: {r}")
    return out[: n.value].copy()


def decode(stream, raw_size_hint: int) -> np.ndarray:
    a = _u8(stream)
    out = np.empty(max(raw_size_hint, 16) + 64, dtype=np.uint8)
    n = C.c_size_t(0)
    r = lib().ref_decode(a.ctypes.data, a.size, out.ctypes.data, out.size, C.byref(n))
    if r < 0:
        raise RuntimeError(f"ref_decode failed: {r}")
    return out[: n.value].copy()


def props(dict_size: int, level: int) -> dict:
    o = np.zeros(15, dtype=np.uint64)
    lib().ref_props_init(dict_size, level, o.ctypes.data)
    k = ["dict_size", "csc_blocksize", "raw_blocksize", "hash_bits", "hash_width", "bt_hash_bits", "bt_size",
         "bt_cyc", "good_len", "lz_mode", "DLTFilter", "TXTFilter", "EXEFilter", "est_mem", "sizeof"]
    return {kk: int(v) for kk, v in zip(k, o)}


def analyze_blocks(data):
    a = _u8(data)
    nb = (a.size + 8191) // 8192
    t = np.zeros(nb, dtype=np.uint32)
    b = np.zeros(nb, dtype=np.uint32)
    lib().ref_analyze_blocks(a.ctypes.data, a.size, t.ctypes.data, b.ctypes.data)
    return t, b


def dlt_bpb(data, chn: int) -> int:
    a = _u8(data)
    return int(lib().ref_dlt_bpb(a.ctypes.data, a.size, chn))


def filt(which: int, inverse: bool, data, chn: int = 0):
    a = _u8(data).copy()
    r = lib().ref_filter(which, int(inverse), a.ctypes.data, a.size, chn)
    return a, r


def run_chunks(chunks, dict_size: int, level: int, nthreads: int, decode_mode: bool = False, read_size: int = 0,
               raw_sizes=None):
    """One reference handle per chunk on ``nthreads`` host threads (csarc -p/-t model)."""
    n = len(chunks)
    ins = [_u8(c) for c in chunks]
    if decode_mode:
        caps = [int(s) + 64 for s in raw_sizes]
    else:
        caps = [out_bound(c.size) for c in ins]
    outs = [np.empty(c, dtype=np.uint8) for c in caps]
    ip = (C.c_void_p * n)(*[c.ctypes.data for c in ins])
    isz = (C.c_size_t * n)(*[c.size for c in ins])
    op = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    ocap = (C.c_size_t * n)(*caps)
    on = (C.c_size_t * n)()
    r = lib().ref_run_chunks(int(decode_mode), n, ip, isz, op, ocap, on, dict_size, level, read_size, nthreads)
    if r < 0:
        raise RuntimeError(f"ref_run_chunks failed: {r}")
    return [outs[i][: on[i]] for i in range(n)]
