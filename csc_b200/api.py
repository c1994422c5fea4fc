"""ctypes binding of the libcsc C ABI served by csrc/libcsc_b200.so.

Mirrors the reference interface for this path: ``props_init`` is
CSCEncProps_Init (csc_enc.cpp:16), ``encode_stream`` / ``decode_stream`` drive
one CSCEnc / CSCDec handle through ISeqInStream / ISeqOutStream callbacks
exactly as csa_worker.cpp:23-90 does, ``encode_chunks`` / ``decode_chunks`` are
the batched calls of include/csc_b200.h (many independent handles at once, the
csarc -p/-t split).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("CSC_B200_LIB") or os.path.join(_HERE, "csrc", "libcsc_b200.so")   # override: A/B builds only

EXPORTS = ["CSCEncProps_Init", "CSCEnc_WriteProperties", "CSCEnc_EstMemUsage", "CSCEnc_Create", "CSCEnc_Destroy",
           "CSCEnc_Encode", "CSCEnc_Encode_Flush", "CSCDec_ReadProperties", "CSCDec_Create", "CSCDec_Destroy",
           "CSCDec_Decode", "CSCB200_EncodeBound", "CSCB200_EncodeChunks", "CSCB200_DecodeChunks",
           "CSCB200_EncodeChunksDevice", "CSCB200_SetDevice", "CSCB200_SetDevices", "CSCB200_DeviceCount", "CSCB200_Trim", "CSCB200_ChunkRange",
           "CSCB200_GetStats", "CSCB200_Available", "CSCB200_EncodeChunksAdler", "CSCB200_DecodeChunksAdler", "CSCB200_Adler32"]


class CSCProps(C.Structure):   # csc_common.h:19-63 (ABI)
    _fields_ = [("dict_size", C.c_size_t), ("csc_blocksize", C.c_uint32), ("raw_blocksize", C.c_uint32),
                ("hash_bits", C.c_uint8), ("hash_width", C.c_uint8), ("bt_hash_bits", C.c_uint8),
                ("bt_size", C.c_uint32), ("bt_cyc", C.c_uint32), ("good_len", C.c_uint8), ("lz_mode", C.c_uint8),
                ("DLTFilter", C.c_uint8), ("TXTFilter", C.c_uint8), ("EXEFilter", C.c_uint8)]


class Stats(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("ms_h2d", "ms_analyze", "ms_filters", "ms_mf_hash", "ms_mf_sort", "ms_mf_rows",
                                          "ms_mf_extend", "ms_encode", "ms_d2h", "ms_frame", "ms_decode", "ms_unfilter",
                                          "ms_total_gpu", "ms_mf_chains", "ms_rc")] + \
               [(k, C.c_uint64) for k in ("bytes_in", "bytes_out", "mf_positions")] + \
               [(k, C.c_uint32) for k in ("kernel_launches", "streams", "waves", "replans")]


READ_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t))
WRITE_FN = C.CFUNCTYPE(C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t)


class _InStream(C.Structure):
    _fields_ = [("Read", READ_FN)]


class _OutStream(C.Structure):
    _fields_ = [("Write", WRITE_FN)]


_lib = None


def lib():
    """Loads the CUDA library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError(f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(csc_b200 has no CPU fallback)")
        l = C.CDLL(_SO)
        l.CSCEncProps_Init.argtypes = [C.POINTER(CSCProps), C.c_uint32, C.c_int]
        l.CSCEnc_WriteProperties.argtypes = [C.POINTER(CSCProps), C.c_void_p, C.c_int]
        l.CSCEnc_EstMemUsage.argtypes = [C.POINTER(CSCProps)]
        l.CSCEnc_EstMemUsage.restype = C.c_uint64
        l.CSCEnc_Create.argtypes = [C.POINTER(CSCProps), C.c_void_p, C.c_void_p]
        l.CSCEnc_Create.restype = C.c_void_p
        l.CSCEnc_Destroy.argtypes = [C.c_void_p]
        l.CSCEnc_Encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.CSCEnc_Encode_Flush.argtypes = [C.c_void_p]
        l.CSCDec_ReadProperties.argtypes = [C.POINTER(CSCProps), C.c_void_p]
        l.CSCDec_Create.argtypes = [C.POINTER(CSCProps), C.c_void_p, C.c_void_p]
        l.CSCDec_Create.restype = C.c_void_p
        l.CSCDec_Destroy.argtypes = [C.c_void_p]
        l.CSCDec_Decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.CSCB200_EncodeBound.argtypes = [C.c_size_t]
        l.CSCB200_EncodeBound.restype = C.c_size_t
        l.CSCB200_EncodeChunks.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p]
        l.CSCB200_EncodeChunksDevice.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                                 C.c_void_p]
        l.CSCB200_DecodeChunks.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.CSCB200_EncodeChunksAdler.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.CSCB200_DecodeChunksAdler.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.CSCB200_Adler32.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.CSCB200_SetDevice.argtypes = [C.c_int]
        l.CSCB200_GetStats.argtypes = [C.POINTER(Stats)]
        _lib = l
    return _lib


def available() -> bool:
    return bool(lib().CSCB200_Available())


def set_device(dev: int) -> None:
    r = lib().CSCB200_SetDevice(dev)
    if r < 0:
        raise RuntimeError(f"CSCB200_SetDevice({dev}) failed: {r}")


def set_devices(devs) -> None:
    """this thread's batched calls fan their chunks out over these GPUs (contiguous chunk ranges, no collective)"""
    devs = list(devs)
    arr = (C.c_int * max(len(devs), 1))(*devs)
    r = lib().CSCB200_SetDevices(arr, len(devs))
    if r < 0:
        raise RuntimeError(f"CSCB200_SetDevices({devs}) failed: {r}")


def chunk_range(n_chunks: int, rank: int, world: int):
    """chunks [lo, hi) of part ``rank`` of ``world`` -- the library's own partition rule (no GPU needed)"""
    lo, hi = C.c_int(0), C.c_int(0)
    lib().CSCB200_ChunkRange(n_chunks, rank, world, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def device_count() -> int:
    return int(lib().CSCB200_DeviceCount())


def get_stats() -> dict:
    s = Stats()
    lib().CSCB200_GetStats(C.byref(s))
    return {k: getattr(s, k) for k, _ in Stats._fields_}


def props_init(dict_size: int, level: int) -> CSCProps:
    p = CSCProps()
    lib().CSCEncProps_Init(C.byref(p), dict_size, level)
    return p


def props_dict(p: CSCProps) -> dict:
    d = {k: int(getattr(p, k)) for k, _ in CSCProps._fields_}
    d["est_mem"] = int(lib().CSCEnc_EstMemUsage(C.byref(p)))
    d["sizeof"] = C.sizeof(CSCProps)
    return d


def encode_bound(n: int) -> int:
    return int(lib().CSCB200_EncodeBound(n))


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray, memoryview)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


def split_chunks(data, chunk_size: int):
    """csarc -p style split of one buffer into independent chunks (archiver/csarc.cpp:532-543)."""
    a = _u8(data)
    return [a[o:o + chunk_size] for o in range(0, max(a.size, 1), chunk_size)] if a.size else [a]


def encode_chunks_raw(in_ptrs, in_sizes, out_ptrs, out_caps, dict_size, level, flt_mask=7, with_props=True):
    """Pointer-level batched encode (host buffers; used by bench.py with pinned memory)."""
    n = len(in_ptrs)
    ip = (C.c_void_p * n)(*in_ptrs)
    isz = (C.c_size_t * n)(*in_sizes)
    op = (C.c_void_p * n)(*out_ptrs)
    ocap = (C.c_size_t * n)(*out_caps)
    on = (C.c_size_t * n)()
    r = lib().CSCB200_EncodeChunks(n, ip, isz, dict_size, level, flt_mask, int(with_props), op, ocap, on)
    if r < 0:
        raise RuntimeError(f"CSCB200_EncodeChunks failed: {r}")
    return [int(v) for v in on]


def encode_chunks_device(dev_ptrs, in_sizes, dict_size, level, flt_mask=7, with_props=True):
    """Kernel-only path: inputs already in HBM, outputs stay in HBM; returns framed sizes."""
    n = len(dev_ptrs)
    ip = (C.c_void_p * n)(*dev_ptrs)
    isz = (C.c_size_t * n)(*in_sizes)
    on = (C.c_size_t * n)()
    r = lib().CSCB200_EncodeChunksDevice(n, ip, isz, dict_size, level, flt_mask, int(with_props), on)
    if r < 0:
        raise RuntimeError(f"CSCB200_EncodeChunksDevice failed: {r}")
    return [int(v) for v in on]


def decode_chunks_raw(in_ptrs, in_sizes, out_ptrs, out_caps):
    n = len(in_ptrs)
    ip = (C.c_void_p * n)(*in_ptrs)
    isz = (C.c_size_t * n)(*in_sizes)
    op = (C.c_void_p * n)(*out_ptrs)
    ocap = (C.c_size_t * n)(*out_caps)
    on = (C.c_size_t * n)()
    r = lib().CSCB200_DecodeChunks(n, ip, isz, op, ocap, on)
    if r < 0:
        raise RuntimeError(f"CSCB200_DecodeChunks failed: {r}")
    return [int(v) for v in on]


def adler32(bufs, init=None):
    """csarc's checksum (archiver/csa_adler32.cpp) of every buffer, on the GPU; init: running values to continue from"""
    arrs = [_u8(b) for b in bufs]
    n = len(arrs)
    ip = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    isz = (C.c_size_t * n)(*[a.size for a in arrs])
    ini = (C.c_uint32 * n)(*init) if init is not None else None
    out = (C.c_uint32 * n)()
    r = lib().CSCB200_Adler32(n, ip, isz, ini, out)
    if r < 0:
        raise RuntimeError(f"CSCB200_Adler32 failed: {r}")
    return [int(v) for v in out]


def encode_chunks_adler(chunks, dict_size: int, level: int, flt_mask: int = 7, with_props: bool = True):
    """encode_chunks plus the Adler-32 of every raw chunk (taken from its copy in HBM)"""
    ins = [_u8(c) for c in chunks]
    n = len(ins)
    outs = [np.empty(encode_bound(c.size), dtype=np.uint8) for c in ins]
    ip = (C.c_void_p * n)(*[c.ctypes.data for c in ins])
    isz = (C.c_size_t * n)(*[c.size for c in ins])
    op = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    ocap = (C.c_size_t * n)(*[o.size for o in outs])
    on = (C.c_size_t * n)()
    ad = (C.c_uint32 * n)()
    r = lib().CSCB200_EncodeChunksAdler(n, ip, isz, dict_size, level, flt_mask, int(with_props), op, ocap, on, ad)
    if r < 0:
        raise RuntimeError(f"CSCB200_EncodeChunksAdler failed: {r}")
    return [o[:int(k)] for o, k in zip(outs, on)], [int(v) for v in ad]


def decode_chunks_adler(streams, raw_sizes):
    """decode_chunks plus the Adler-32 of every decoded chunk (taken on the device before the download)"""
    ins = [_u8(c) for c in streams]
    n = len(ins)
    outs = [np.empty(int(k) + 64, dtype=np.uint8) for k in raw_sizes]
    ip = (C.c_void_p * n)(*[c.ctypes.data for c in ins])
    isz = (C.c_size_t * n)(*[c.size for c in ins])
    op = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    ocap = (C.c_size_t * n)(*[int(k) for k in raw_sizes])
    on = (C.c_size_t * n)()
    ad = (C.c_uint32 * n)()
    r = lib().CSCB200_DecodeChunksAdler(n, ip, isz, op, ocap, on, ad)
    if r < 0:
        raise RuntimeError(f"CSCB200_DecodeChunksAdler failed: {r}")
    return [o[:int(k)] for o, k in zip(outs, on)], [int(v) for v in ad]


def encode_chunks(chunks, dict_size: int, level: int, flt_mask: int = 7, with_props: bool = True):
    """Each chunk -> exactly the stream one CSCEnc handle writes for it."""
    ins = [_u8(c) for c in chunks]
    outs = [np.empty(encode_bound(c.size), dtype=np.uint8) for c in ins]
    sizes = encode_chunks_raw([c.ctypes.data for c in ins], [c.size for c in ins], [o.ctypes.data for o in outs],
                              [o.size for o in outs], dict_size, level, flt_mask, with_props)
    return [o[:s] for o, s in zip(outs, sizes)]


def decode_chunks(streams, raw_sizes):
    ins = [_u8(c) for c in streams]
    outs = [np.empty(int(s) + 64, dtype=np.uint8) for s in raw_sizes]
    sizes = decode_chunks_raw([c.ctypes.data for c in ins], [c.size for c in ins], [o.ctypes.data for o in outs],
                              [int(s) for s in raw_sizes])
    return [o[:s] for o, s in zip(outs, sizes)]


def encode_traced(data, dict_size: int, level: int):
    """Debug (tests): one chunk through CSCB200_EncodeChunks plus the engine's trace as (kind, a, b) triples --
    packets 0..3 and, per find_match call, (4, position, n) followed by n x (5, len, dist code)."""
    a = _u8(data)
    l = lib()
    l.CSCB200_DebugTrace.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    tr = np.zeros(((a.size + 1024) * 12, 3), dtype=np.uint32)
    n = C.c_size_t(0)
    l.CSCB200_DebugTrace(tr.ctypes.data, tr.shape[0], C.byref(n))
    try:
        out = encode_chunks([a], dict_size, level)[0]
    finally:
        l.CSCB200_DebugTrace(None, 0, None)
    if n.value > tr.shape[0]:
        raise RuntimeError("trace buffer too small")
    return out, tr[: n.value]


def encode_stream(data, dict_size: int, level: int, read_size: int = 0, flt_mask: int = 7, with_props: bool = True):
    """One CSCEnc handle over ISeqInStream/ISeqOutStream callbacks (the csa_worker.cpp:35-50 sequence).
    ``dict_size`` is the argument of CSCEncProps_Init; ``read_size`` caps the bytes returned per Read()."""
    a = _u8(data)
    l = lib()
    p = props_init(dict_size, level)
    p.DLTFilter, p.TXTFilter, p.EXEFilter = flt_mask & 1, (flt_mask >> 1) & 1, (flt_mask >> 2) & 1
    out = bytearray()
    if with_props:
        hdr = (C.c_uint8 * 10)()
        l.CSCEnc_WriteProperties(C.byref(p), hdr, 0)
        out += bytes(hdr)
    state = {"off": 0}

    def rd(_p, buf, psize):
        want = psize[0]
        if read_size:
            want = min(want, read_size)
        want = min(want, a.size - state["off"])
        if want:
            C.memmove(buf, a.ctypes.data + state["off"], want)
        state["off"] += want
        psize[0] = want
        return 0

    def wr(_p, buf, size):
        out.extend(C.string_at(buf, size))
        return size

    ins = _InStream(READ_FN(rd))
    outs = _OutStream(WRITE_FN(wr))
    h = l.CSCEnc_Create(C.byref(p), C.byref(outs), None)
    if not h:
        raise RuntimeError("CSCEnc_Create failed")
    try:
        r = l.CSCEnc_Encode(h, C.byref(ins), None)
        if r < 0:
            raise RuntimeError(f"CSCEnc_Encode failed: {r}")
        r = l.CSCEnc_Encode_Flush(h)
        if r < 0:
            raise RuntimeError(f"CSCEnc_Encode_Flush failed: {r}")
    finally:
        l.CSCEnc_Destroy(h)
    return np.frombuffer(bytes(out), dtype=np.uint8)


def decode_stream(stream):
    """One CSCDec handle (csa_worker.cpp:75-83): the caller parses the 10-byte header."""
    a = _u8(stream)
    l = lib()
    p = CSCProps()
    hdr = (C.c_uint8 * 10)(*a[:10].tolist())
    l.CSCDec_ReadProperties(C.byref(p), hdr)
    state = {"off": 10}
    out = bytearray()

    def rd(_p, buf, psize):
        want = min(psize[0], a.size - state["off"])
        if want:
            C.memmove(buf, a.ctypes.data + state["off"], want)
        state["off"] += want
        psize[0] = want
        return 0

    def wr(_p, buf, size):
        out.extend(C.string_at(buf, size))
        return size

    ins = _InStream(READ_FN(rd))
    outs = _OutStream(WRITE_FN(wr))
    h = l.CSCDec_Create(C.byref(p), C.byref(ins), None)
    if not h:
        raise RuntimeError("CSCDec_Create failed")
    try:
        r = l.CSCDec_Decode(h, C.byref(outs), None)
        if r < 0:
            raise RuntimeError(f"CSCDec_Decode failed: {r}")
    finally:
        l.CSCDec_Destroy(h)
    return np.frombuffer(bytes(out), dtype=np.uint8)
