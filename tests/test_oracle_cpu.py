"""CPU-only tests (-m "not gpu"): the oracle against the reference, the host-side logic of the
product (props, framing, typing, filters, engines in single-lane host simulation) against the
oracle, and that the C-ABI library loads and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from csc_b200 import synth

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
EC = synth.edge_cases()


def same(a, b):
    return a.size == b.size and bool((a == b).all())


# ---------------------------------------------------------------- boundary
def test_library_exports_all_declared_symbols():
    import csc_b200
    l = csc_b200.lib()
    declared = set()
    for h in ("csc_enc.h", "csc_dec.h", "csc_b200.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        declared |= set(re.findall(r"\b(CSC(?:Enc|Dec|EncProps|B200)_\w+)\s*\(", txt))
    assert declared >= set(csc_b200.EXPORTS) - {"CSCB200_DebugDump"}
    for name in declared:
        assert hasattr(l, name), f"{name} declared in include/ but not exported"


def test_cscprops_abi_layout():
    import csc_b200
    # csc_common.h:19-63: size_t, u32, u32, u8 x3, u32, u32, u8 x5 -> 40 bytes on LP64
    assert ctypes.sizeof(csc_b200.CSCProps) == 40
    assert csc_b200.CSCProps.dict_size.offset == 0 and csc_b200.CSCProps.csc_blocksize.offset == 8
    assert csc_b200.CSCProps.bt_size.offset == 20 and csc_b200.CSCProps.good_len.offset == 28


@pytest.mark.parametrize("level", [0, 1, 2, 3, 4, 5, 9])
@pytest.mark.parametrize("dict_size", [1000, 32 << 10, 64 << 10, 256 << 10, 1 << 20, (1 << 20) + 4, 4 << 20, 16 << 20,
                                       32 << 20, 64 << 20, 256 << 20, 1000 << 20, 0xFFFFFFFF])
def test_props_init_matches_reference(ref, level, dict_size):
    import csc_b200
    from csc_b200 import api
    exp = ref.props(dict_size, level)
    got = api.props_dict(csc_b200.props_init(dict_size, level))
    for k in exp:
        assert got[k] == exp[k], (k, got[k], exp[k])


def test_props_header_roundtrip():
    import csc_b200
    l = csc_b200.lib()
    p = csc_b200.props_init(5 << 20, 3)
    hdr = (ctypes.c_uint8 * 10)()
    l.CSCEnc_WriteProperties(ctypes.byref(p), hdr, 0)
    q = csc_b200.CSCProps()
    l.CSCDec_ReadProperties(ctypes.byref(q), hdr)
    assert (q.dict_size, q.csc_blocksize, q.raw_blocksize) == (p.dict_size, 65536, 2 << 20)


def test_no_cpu_fallback_without_gpu():
    import csc_b200
    if csc_b200.available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        csc_b200.encode_chunks([EC["tiny"]], 1 << 20, 1)


# ---------------------------------------------------------------- class level
@pytest.mark.parametrize("name", ["mixed_small", "skip_tail", "entropy", "lowent", "delta1", "delta3", "delta8", "zeros",
                                  "text_hibytes", "exe_small"])
def test_analyzer_matches_reference(ref, hs, name):
    d = EC[name]
    rt, rb = ref.analyze_blocks(d)
    ht, hb = hs.analyze_blocks(d)
    assert ((ht & 0xFF) == rt).all()
    m = rt != 0x1E   # bpb is not written for DT_SKIP blocks
    assert (hb[m] == rb[m]).all()
    # the dlt veto bit equals GetDltBpb >= bpb * 0.95 of the reference
    for i in np.nonzero(rt >= 0x10)[0]:
        if rt[i] == 0x1E:
            continue
        blk = d[i * 8192:(i + 1) * 8192]
        chn = [1, 2, 3, 4, 8][rt[i] - 0x10]
        assert ((ht[i] >> 8) & 1) == int(ref.dlt_bpb(blk, chn) >= rb[i] * 0.95)


def test_analyzer_reaches_every_type(ref):
    seen = set()
    for v in EC.values():
        if v.size:
            seen |= set(ref.analyze_blocks(v)[0].tolist())
    assert {1, 2, 3, 7, 8, 0x10, 0x11, 0x12, 0x13, 0x14, 0x1E} <= seen, seen


@pytest.mark.parametrize("which,name,chn", [(0, "exe_small", 0), (0, "mixed_small", 0), (1, "text_small", 0),
                                            (1, "text_hibytes", 0), (1, "entropy", 0), (2, "delta3", 3), (2, "delta8", 8),
                                            (2, "tiny", 2)])
def test_filters_match_reference(ref, hs, which, name, chn):
    d = EC[name]
    rf, rok = ref.filt(which, False, d, chn)
    hf, hok = hs.filt(which, False, d, chn)
    assert rok == hok and same(rf, hf)
    ri, _ = ref.filt(which, True, rf, chn)
    hi, _ = hs.filt(which, True, rf, chn)
    assert same(ri, hi)
    if rok:
        assert same(hi, d)


def test_e8e9_pathological_opcode_runs(ref, hs):
    rng = np.random.default_rng(3)
    d = rng.choice(np.array([0xE8, 0xE9, 0x00, 0xFF, 0x12], np.uint8), size=20000)
    rf, _ = ref.filt(0, False, d)
    hf, _ = hs.filt(0, False, d)
    assert same(rf, hf)
    assert same(hs.filt(0, True, hf)[0], d)


# ---------------------------------------------------------------- stream level
@pytest.mark.parametrize("level", [1, 2, 3, 4])
@pytest.mark.parametrize("name", list(EC.keys()))
def test_hostsim_stream_identical_to_reference(ref, hs, name, level):
    d = EC[name]
    if name == "mixed_small" and level in (2, 4):
        d = d[:400000]   # keeps the brute-force K=8 row builder of the simulation fast
    dict_size = max(d.size, 32768)
    assert same(hs.encode(d, dict_size, level), ref.encode(d, dict_size, level))


def test_hostsim_m5_bounded_search_ratio_and_reference_decode(ref, hs):
    """-m5 is ratio-gated (BASELINE north_star): the bounded multi-context search must decode bit-exact with
    the reference decoder and stay within 0.5 % of libcsc -m5's size."""
    for name, d in {"text": synth.text(400000, 11), "mixed": EC["mixed_small"][:600000],
                    "redundant": synth.redundant(700000, base=250000, seed=12)}.items():
        ours, theirs = hs.encode(d, d.size, 5), ref.encode(d, d.size, 5)
        assert same(ref.decode(ours, d.size), d), name
        assert same(hs.decode(ours, d.size), d), name
        assert ours.size <= theirs.size * 1.005, (name, ours.size, theirs.size)


@pytest.mark.parametrize("read_size", [8192, 100000, 300007])
def test_hostsim_read_granularity_is_part_of_the_stream(ref, hs, read_size):
    d = synth.mixed(700000, 8)
    for level in (1, 3):
        assert same(hs.encode(d, d.size, level, read_size), ref.encode(d, d.size, level, read_size))


@pytest.mark.parametrize("flt", [0, 1, 2, 4, 6])
def test_hostsim_filter_switches(ref, hs, flt):
    d = synth.mixed(600000, 9)
    assert same(hs.encode(d, d.size, 1, flt_mask=flt), ref.encode(d, d.size, 1, flt_mask=flt))


def test_hostsim_rc_block_boundaries(ref, hs):
    # > 64 KiB of range-coder and of raw-bit output in one segment: block interleaving (H4) + stale byte (H3)
    d = np.concatenate([synth.random_bytes(150000, 1), synth.text(200000, 2), synth.exe(300000, 3)])
    for level in (1, 3):
        assert same(hs.encode(d, d.size, level), ref.encode(d, d.size, level))


@pytest.mark.parametrize("level", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("name", ["empty", "one", "zeros", "skip_tail", "entropy", "delta3", "dupblock", "text_hibytes",
                                  "mixed_small", "longmatch"])
def test_hostsim_decoder_reads_reference_streams(ref, hs, name, level):
    d = EC[name]
    s = ref.encode(d, max(d.size, 32768), level)
    assert same(hs.decode(s, d.size), d)


def test_hostsim_decoder_handles_window_wrap(ref, hs):
    d = synth.mixed(300000, 17)
    s = ref.encode(d, 40000, 2)   # dictionary much smaller than the stream: the reference window wraps
    assert same(ref.decode(s, d.size), d)
    assert same(hs.decode(s, d.size), d)


def test_golden_fixtures(hs):
    """Golden vectors generated from the unmodified reference (tests/golden/make_golden.py)."""
    import hashlib
    import json
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
    for rec in [r for r in g["streams"] if r["level"] <= 4]:
        d = EC[rec["name"]]
        assert hashlib.sha256(d.tobytes()).hexdigest() == rec["input_sha256"], "synthetic generator drifted"
        out = hs.encode(d, max(d.size, 32768), rec["level"])
        assert out.size == rec["size"] and hashlib.sha256(out.tobytes()).hexdigest() == rec["sha256"], rec
    for rec in g["small_streams"]:
        d = np.frombuffer(bytes.fromhex(rec["input_hex"]), np.uint8)
        exp = np.frombuffer(bytes.fromhex(rec["stream_hex"]), np.uint8)
        assert same(hs.encode(d, 32768, rec["level"]), exp)
        assert same(hs.decode(exp, d.size), d)


# ---------------------------------------------------------------- the CPU restatement (oracle/restate)
@pytest.mark.parametrize("level", [1, 2, 3, 4])
def test_restatement_pinned_to_reference(ref, level):
    from oracle import restate
    for name, d in EC.items():
        dict_size = max(d.size, 32768)
        exp = ref.encode(d, dict_size, level)
        assert same(restate.encode(d, dict_size, level), exp), name
        assert same(restate.decode(exp, d.size), d), name


def test_restatement_window_wrap_and_segments(ref):
    from oracle import restate
    d = synth.mixed(700000, 8)
    for level in (1, 3):
        for dict_size in (40000, 200000):
            for rs in (0, 100000):
                exp = ref.encode(d, dict_size, level, rs)
                assert same(restate.encode(d, dict_size, level, rs), exp)
                assert same(restate.decode(exp, d.size), d)


def test_restatement_matches_golden():
    """pins the restatement without the reference (the GPU box has no /root/reference)"""
    import hashlib
    import json
    from oracle import restate
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
    for rec in [r for r in g["streams"] if r["level"] <= 4]:
        d = EC[rec["name"]]
        out = restate.encode(d, max(d.size, 32768), rec["level"])
        assert out.size == rec["size"] and hashlib.sha256(out.tobytes()).hexdigest() == rec["sha256"], rec


# ---------------------------------------------------------------- drop-in boundary
def test_reference_cli_builds_against_product_headers_and_library(ref, tmp_path):
    """the reference's own tool (libcsc/csc.cpp, unmodified) compiles against
    include/ and links with libcsc_b200.so.  Without a CUDA device the relinked tool must fail loudly and
    produce nothing but the caller-written 10-byte header: there is no CPU fallback."""
    import subprocess
    try:
        exe, exe_ref = ref.cli("b200"), ref.cli("ref")
    except FileNotFoundError:
        pytest.skip("oracle/_ref/csc_cli_* not built and /root/reference absent")
    ldd = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libcsc_b200.so" in ldd and "not found" not in ldd
    d = synth.mixed(300000, 21)
    src, out, out_ref = tmp_path / "in.bin", tmp_path / "out.csc", tmp_path / "ref.csc"
    d.tofile(src)
    r = subprocess.run([exe_ref, "c", "-m1", str(src), str(out_ref)], capture_output=True, text=True)
    assert r.returncode == 0
    assert same(np.fromfile(out_ref, dtype=np.uint8), ref.encode(d, d.size, 1))   # the tool == the library oracle
    import csc_b200
    csc_b200.lib()
    if not csc_b200.available():
        r = subprocess.run([exe, "c", "-m1", str(src), str(out)], capture_output=True, text=True)
        assert "no CPU fallback" in r.stderr
        assert os.path.getsize(out) == 10


def test_reference_archiver_links_against_the_product(ref, tmp_path):
    """csarc (archiver/*.cpp, unmodified) builds against include/ + libcsc_b200.so.  The reference build of the
    same sources round-trips here; without a CUDA device the relinked one reports the missing device and
    terminates (its producer threads are drained) instead of hanging."""
    import subprocess
    try:
        exe, exe_ref = ref.cli("csarc_b200"), ref.cli("csarc_ref")
    except FileNotFoundError:
        pytest.skip("oracle/_ref/csarc_* not built and /root/reference absent")
    ldd = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libcsc_b200.so" in ldd and "not found" not in ldd
    files = {"a.txt": synth.text(200000, 31), "b.exe": synth.exe(150000, 32)}
    for n, d in files.items():
        d.tofile(tmp_path / n)
    names = list(files)
    r = subprocess.run([exe_ref, "a", "-m2", "-t2", "-f", "ref.csa"] + names, cwd=tmp_path, capture_output=True, timeout=120)
    assert r.returncode == 0
    os.mkdir(tmp_path / "out")
    r = subprocess.run([exe_ref, "x", "-t2", "-o", "out/", "ref.csa"], cwd=tmp_path, capture_output=True, timeout=120)
    assert r.returncode == 0
    for n, d in files.items():
        assert same(np.fromfile(tmp_path / "out" / n, dtype=np.uint8), d)
    import csc_b200
    csc_b200.lib()
    if not csc_b200.available():
        r = subprocess.run([exe, "a", "-m2", "-t2", "-f", "gpu.csa"] + names, cwd=tmp_path, capture_output=True, text=True, timeout=60)
        assert "no CPU fallback" in r.stderr


# ---------------------------------------------------------------- match finder, class level
# north_star: "per-position match candidates bit-exact against the reference MatchFinder".  The reference
# side is the UNMODIFIED MatchFinder with its two entry points intercepted at link time
# (oracle/ref_mftrace.cpp); ours is the engine's candidate list of every find_match call.  The traces
# interleave packets (0..3) with candidate records (4: position + count, 5: len + dist code), so equality
# also pins WHERE along the parse each search happened.
@pytest.mark.parametrize("level", [1, 2, 3, 4])
@pytest.mark.parametrize("name", ["mixed_small", "text_small", "exe_small", "longmatch", "zero_runs", "dupblock"])
def test_match_candidates_equal_reference_matchfinder(ref, hs, level, name):
    d = EC[name]
    rs, rt = ref.encode(d, d.size, level, cands=True)
    hs_, ht = hs.encode(d, d.size, level, cands=True)
    assert same(rs, hs_)
    assert rt.shape == ht.shape and (rt == ht).all(), first_diff(rt, ht)
    assert int((rt[:, 0] == 4).sum()) > 0


def first_diff(a, b):
    n = min(len(a), len(b))
    bad = np.nonzero((a[:n] != b[:n]).any(axis=1))[0]
    i = int(bad[0]) if bad.size else n
    return f"first differing trace record {i}: reference {a[i:i + 3].tolist()} vs ours {b[i:i + 3].tolist()}"


# ---------------------------------------------------------------- window wrap (streams longer than their dictionary)
# LZ::EncodeNormal re-grids the <= 8 KiB sub-blocks at the window end (csc_lz.cpp:65-66,91-92), tail hashes
# read the previous lap's bytes, compares stop at the window end on the candidate's side
# (csc_mf.cpp:268,307,338,460), and the optimal parser's literal context is dropped at window offset 0
# (csc_lz.cpp:297).  Streams AND per-call candidate lists must equal the reference's.
WRAP_CASES = [("mixed600k", lambda: synth.mixed(600000, 17), 40000, 0), ("mixed1M", lambda: EC["mixed_small"], 100000, 0),
              ("text", lambda: synth.text(300000, 3), 33000, 0), ("dupblock", lambda: EC["dupblock"], 20000, 0),
              ("dupdelta", lambda: EC["dupdelta"], 30000, 0), ("longmatch", lambda: EC["longmatch"], 50000, 0),
              ("zero_runs", lambda: EC["zero_runs"], 40000, 0), ("exe_odd_reads", lambda: synth.exe(400000, 5), 70001, 30011)]


@pytest.mark.parametrize("level", [1, 2, 3, 4])
@pytest.mark.parametrize("case", WRAP_CASES, ids=[c[0] for c in WRAP_CASES])
def test_window_wrap_matches_reference(ref, hs, case, level):
    _, gen, dict_size, read_size = case
    d = gen()
    assert d.size > dict_size + 10240
    rs, rt = ref.encode(d, dict_size, level, read_size, cands=True)
    hs.lib().hs_replans()
    hs_, ht = hs.encode(d, dict_size, level, read_size, cands=True)
    assert same(rs, hs_)
    assert rt.shape == ht.shape and (rt == ht).all(), first_diff(rt, ht)
    assert same(hs.decode(hs_, d.size), d)


def test_window_wrap_replans_on_duplicate_block_flips(ref, hs):
    """the static sub-block grid assumes the outcome of IsDuplicateBlock; a wrong guess on a wrapped stream is
    detected by the engine and planned again (plan.h / Encoder::compress_segment)"""
    d = EC["dupblock"]
    hs.lib().hs_replans()
    assert same(hs.encode(d, 20000, 2), ref.encode(d, 20000, 2))
    assert hs.lib().hs_replans() >= 1


def test_device_framing_statement_equals_framer(hs):
    """frame.cuh (k_frame on the device) against host_common.h: Framer on synthetic sub-streams: segments with and without
    full RC blocks, remainders longer and shorter than the ones before (every branch of the stale-byte search),
    random interleaving of RC / BC block completions (H4), small and default block sizes"""
    l = hs.lib()
    l.hs_frame_selftest.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
    for seed in range(300):
        for nseg, bsize in ((1, 64), (2, 64), (5, 64), (9, 256), (3, 65536)):
            if bsize == 65536 and seed >= 12:
                continue
            assert l.hs_frame_selftest(seed, nseg, bsize) == 0, (seed, nseg, bsize)
