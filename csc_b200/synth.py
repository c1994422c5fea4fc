"""Deterministic synthetic corpora for parity tests and bench.py.

The recipes follow SURVEY.md section 8(d) / BASELINE.md section 3 (numpy
``default_rng``, base seed 12345, tile *i* uses seed 12345+i).  They are the
workloads BASELINE.json's configs name: English-like *text*, x86-like *exe*
(for the e8e9 filter), *delta* tables (2/4 channels), *random* bytes, the
*mixed* corpus and the long-distance *redundant* corpus.
"""
from __future__ import annotations

import numpy as np

_WORDS = (
    "the of and to in a is that for it as was with be by on not he i this are or his from at "
    "which but have an had they you were their one all we can her has there been if more when "
    "will would who so no out up said what its about than into them only other time new some "
    "could these two may first then do any like my now over such our man me even most made "
    "after also did many before must well back through years where much your way down should "
    "because long each just state those people too how little good world make very year still "
    "see own work men day get here between both life being under three never know same last "
    "another while us off might great states go come since against right came take used himself "
    "few house use during without again place american around however home small found mrs "
    "thought went say part once general high upon school every does got united left number "
    "course war until always away something fact water though less public put think almost hand "
    "enough far took head yet government system better set told nothing night end why called"
).split()
assert len(_WORDS) >= 200
_WORDS = _WORDS[:200]


def text(nbytes: int, seed: int = 12345) -> np.ndarray:
    """English-like text: 200-word vocabulary, Zipf 1/rank, space-joined,
    newline every 20 000 words, ' the of ' -> '. The of '."""
    rng = np.random.default_rng(seed)
    wl = np.array([len(w) for w in _WORDS])
    maxl = int(wl.max())
    table = np.zeros((len(_WORDS), maxl + 1), dtype=np.uint8)
    for i, w in enumerate(_WORDS):
        table[i, : len(w)] = np.frombuffer(w.encode(), dtype=np.uint8)
        table[i, len(w)] = 0x20
    p = 1.0 / np.arange(1, len(_WORDS) + 1)
    p /= p.sum()
    avg = float((p * (wl + 1)).sum())
    out = []
    have = 0
    while have < nbytes:
        nw = 20000
        idx = rng.choice(len(_WORDS), size=nw, p=p)
        rows = table[idx]
        lens = wl[idx] + 1
        mask = np.arange(maxl + 1)[None, :] < lens[:, None]
        buf = rows[mask].copy()
        buf[-1] = 0x0A
        b = buf.tobytes().replace(b" the of ", b". The of ")
        out.append(b)
        have += len(b)
    del avg
    return np.frombuffer(b"".join(out), dtype=np.uint8)[:nbytes].copy()


_OPS = np.array([0x8B, 0x00, 0x00, 0x89, 0x55, 0x48, 0x83, 0xC3, 0xFF, 0x0F, 0x45, 0xEC, 0x24, 0x04],
                dtype=np.uint8)


def exe(nbytes: int, seed: int = 12345) -> np.ndarray:
    """x86-like code: 70 % opcode alphabet / 30 % uniform, one E8 rel32 call
    every ~40 bytes to one of 64 targets (rel = target - (pos+5))."""
    rng = np.random.default_rng(seed)
    a = _OPS[rng.integers(0, len(_OPS), size=nbytes)]
    u = rng.integers(0, 256, size=nbytes, dtype=np.uint8)
    sel = rng.random(nbytes) < 0.7
    buf = np.where(sel, a, u).astype(np.uint8)
    if nbytes > 64:
        ncall = nbytes // 40
        pos = np.sort(rng.choice(nbytes - 5, size=ncall, replace=False)).astype(np.int64)
        targets = rng.integers(0, nbytes, size=64).astype(np.int64)
        tgt = targets[rng.integers(0, 64, size=ncall)]
        rel = (tgt - (pos + 5)).astype(np.int64) & 0xFFFFFFFF
        buf[pos] = 0xE8
        for k in range(4):
            buf[pos + 1 + k] = ((rel >> (8 * k)) & 0xFF).astype(np.uint8)
    return buf


def delta(nbytes: int, ch: int, seed: int = 12345) -> np.ndarray:
    """Interleaved ``ch``-channel tables: per channel cumsum of U{-3..3} + base."""
    rng = np.random.default_rng(seed)
    n = (nbytes + ch - 1) // ch
    steps = rng.integers(-3, 4, size=(n, ch))
    base = rng.integers(0, 256, size=(1, ch))
    v = (np.cumsum(steps, axis=0) + base) & 0xFF
    return v.astype(np.uint8).reshape(-1)[:nbytes].copy()


def random_bytes(nbytes: int, seed: int = 12345) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 256, size=nbytes, dtype=np.uint8)


_MIX = (("text", 0.375), ("exe", 0.1875), ("delta4", 0.0625), ("text", 0.1875), ("delta2", 0.0625),
        ("random", 0.031), ("exe", 0.094))


def mixed(nbytes: int, seed: int = 12345) -> np.ndarray:
    """text 37.5 % | exe 18.75 % | delta-4 6.25 % | text 18.75 % | delta-2 6.25 %
    | random 3.1 % | exe 9.4 %."""
    parts = []
    used = 0
    for k, (kind, frac) in enumerate(_MIX):
        n = nbytes - used if k == len(_MIX) - 1 else int(nbytes * frac)
        s = seed + 101 * k
        if kind == "text":
            parts.append(text(n, s))
        elif kind == "exe":
            parts.append(exe(n, s))
        elif kind == "delta4":
            parts.append(delta(n, 4, s))
        elif kind == "delta2":
            parts.append(delta(n, 2, s))
        else:
            parts.append(random_bytes(n, s))
        used += n
    return np.concatenate(parts)


def tiled(kind: str, nbytes: int, tile: int = 64 << 20, seed: int = 12345) -> np.ndarray:
    """Large corpora as tiles with per-tile seeds 12345+i."""
    fn = {"text": text, "mixed": mixed, "exe": exe, "random": random_bytes}[kind]
    out = np.empty(nbytes, dtype=np.uint8)
    off = 0
    i = 0
    while off < nbytes:
        n = min(tile, nbytes - off)
        out[off:off + n] = fn(n, seed + i)
        off += n
        i += 1
    return out


def redundant(nbytes: int, base: int = 64 << 20, seed: int = 12345) -> np.ndarray:
    """Config-4 style: a text/exe base followed by copies mutated at 1 % of bytes."""
    rng = np.random.default_rng(seed)
    base = min(base, nbytes)
    b = np.concatenate([text(base // 2, seed), exe(base - base // 2, seed + 1)])
    out = np.empty(nbytes, dtype=np.uint8)
    off = 0
    while off < nbytes:
        n = min(base, nbytes - off)
        c = b[:n].copy()
        if off:
            m = rng.choice(n, size=max(1, n // 100), replace=False)
            c[m] = rng.integers(0, 256, size=len(m), dtype=np.uint8)
        out[off:off + n] = c
        off += n
    return out


def edge_cases(seed: int = 7) -> dict:
    """Small buffers that reach every analyzer outcome and the awkward paths:
    SKIP tails, ENTROPY (6..15 symbols), NORMAL, 1/3/8-channel tables, zero
    runs (hash 0 / run de-dup), long matches (>129, SlidePos skip zone)."""
    rng = np.random.default_rng(seed)
    d = {}
    d["empty"] = np.zeros(0, np.uint8)
    d["one"] = np.array([65], np.uint8)
    d["tiny"] = np.frombuffer(b"hello hello hello", np.uint8).copy()
    d["zeros"] = np.zeros(40000, np.uint8)
    d["skip_tail"] = np.concatenate([text(8192 * 3, 3), text(300, 4)])
    d["entropy"] = rng.choice(np.array([3, 17, 40, 99, 150, 200, 250, 7, 9], np.uint8), size=8192 * 5)
    d["lowent"] = rng.choice(np.arange(40, 60, dtype=np.uint8), size=8192 * 4)
    d["delta1"] = delta(8192 * 4, 1, 5)
    # 1-channel table (DT_DLT + 0): small steps with a random re-base every ~6 samples
    N1 = 8192 * 3
    v = np.zeros(N1, dtype=np.int64)
    st = rng.integers(-3, 4, size=N1)
    jump = rng.random(N1) < (1.0 / 6)
    newv = rng.integers(0, 256, size=N1)
    cur = 0
    for i in range(N1):
        cur = int(newv[i]) if jump[i] else (cur + int(st[i])) & 0xFF
        v[i] = cur
    d["delta1b"] = v.astype(np.uint8)
    d["delta3"] = delta(8192 * 4 + 100, 3, 6)
    d["delta8"] = delta(8192 * 6, 8, 7)
    long_rep = np.tile(text(5000, 9), 12)
    d["longmatch"] = np.concatenate([long_rep, random_bytes(3000, 1), long_rep[:30000]])
    zr = np.zeros(100000, np.uint8)
    zr[::977] = 1
    zr[50000:50300] = rng.integers(0, 256, 300, dtype=np.uint8)
    d["zero_runs"] = zr
    d["dupblock"] = np.concatenate([random_bytes(8192 * 3, 11), text(8192, 12), random_bytes(8192 * 3, 11)])
    d["dupdelta"] = np.concatenate([delta(8192 * 4, 4, 13), text(8192 * 2, 14), delta(8192 * 4, 4, 13)])
    d["exe_small"] = exe(8192 * 6 + 77, 15)
    d["text_small"] = text(8192 * 5 + 1, 16)
    d["text_short_run"] = np.concatenate([text(8192, 17), random_bytes(8192, 18), text(8192 * 3, 19)])
    hi = text(8192 * 4, 20).copy()
    hi[rng.choice(len(hi), 1500, replace=False)] = rng.integers(0x82, 256, 1500, dtype=np.uint8)
    d["text_hibytes"] = hi
    d["mixed_small"] = mixed(1 << 20, 21)
    return d
