"""Developer diagnostic for a gpurun call: parity of the CUDA path against the oracle with
stage-level bisection (analyzer / filters / match-finder rows) when a stream differs."""
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import ctypes as C

import csc_b200
from csc_b200 import api, synth
from oracle import refbind as R
from tests.hostsim import hsbind as H


def first_diff(a, b):
    n = min(a.size, b.size)
    d = np.nonzero(a[:n] != b[:n])[0]
    return int(d[0]) if d.size else (n if a.size != b.size else -1)


def stage_diff(d, level):
    dict_size = max(d.size, 32768)
    K = max(1, H.props(dict_size, level)["hash_width"])
    _, hst = H.encode_stages(d, dict_size, level)
    gst = H.Stages(d.size, K)
    l = api.lib()
    l.CSCB200_DebugDump.argtypes = [C.c_void_p] * 9
    l.CSCB200_DebugDump(*gst.ptrs())
    try:
        csc_b200.encode_chunks([d], dict_size, level)
    except Exception as e:  # noqa
        print("   gpu encode raised", e)
    l.CSCB200_DebugDump(*([None] * 9))
    n = d.size
    nblk = (n + 8191) // 8192
    for name, a, b, m in (("binfo", hst.binfo, gst.binfo, nblk), ("bbpb", hst.bbpb, gst.bbpb, nblk),
                          ("dict_ok", hst.dict_ok, gst.dict_ok, nblk), ("lz", hst.lz, gst.lz, n),
                          ("Z2", hst.Z2, gst.Z2, n // 32), ("Z6", hst.Z6, gst.Z6, n // 32), ("P6", hst.P6, gst.P6, n * K),
                          ("L6", hst.L6, gst.L6, n * K), ("tail_dlt", hst.tail_dlt, gst.tail_dlt, 5)):
        fd = first_diff(a[:m], b[:m])
        print(f"   stage {name:8s} {'same' if fd < 0 else 'DIFF@' + str(fd)}", "" if fd < 0 else (a[max(0, fd - 2):fd + 4], b[max(0, fd - 2):fd + 4]))


def main():
    print("available", csc_b200.available())
    bad = 0
    ec = synth.edge_cases()
    levels = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "1,2,3,4".split(","))]
    for lv in levels:
        names = list(ec.keys())
        chunks = [ec[k] for k in names]
        t = time.time()
        try:
            outs = csc_b200.encode_chunks(chunks, 1 << 30, lv)
        except Exception as e:
            print("encode_chunks raised", e)
            outs = [np.zeros(0, np.uint8)] * len(chunks)
        dt = time.time() - t
        for k, c, o in zip(names, chunks, outs):
            r = R.encode(c, min(1 << 30, c.size), lv)
            ok = r.size == o.size and (r == o).all()
            if not ok:
                bad += 1
                print(f"m{lv} {k:16s} n={c.size} ref={r.size} gpu={o.size} DIFF@{first_diff(r, o)}")
                if bad <= 3:
                    stage_diff(c, lv)
        print(f"level {lv}: batch of {len(chunks)} edge cases in {dt:.2f}s, bad so far {bad}")
    # decode of reference streams + of our own
    for lv in (1, 3):
        names = list(ec.keys())
        streams = [R.encode(ec[k], min(1 << 30, ec[k].size), lv) for k in names]
        try:
            outs = csc_b200.decode_chunks(streams, [ec[k].size for k in names])
            for k, o in zip(names, outs):
                if o.size != ec[k].size or not (o == ec[k]).all():
                    bad += 1
                    print(f"decode m{lv} {k} MISMATCH size {o.size} vs {ec[k].size} firstdiff {first_diff(o, ec[k])}")
        except Exception as e:
            bad += 1
            print("decode_chunks raised", e)
    print("decode checks done, bad", bad)
    # handle API
    d = synth.mixed(3 << 20, 5)
    for lv in (1, 3):
        try:
            o = csc_b200.encode_stream(d, d.size, lv)
            r = R.encode(d, d.size, lv)
            ok = o.size == r.size and (o == r).all()
            print(f"handle API m{lv}: {'OK' if ok else 'DIFF@' + str(first_diff(o, r))}")
            bad += not ok
            back = csc_b200.decode_stream(o)
            ok = back.size == d.size and (back == d).all()
            print(f"handle decode m{lv}: {'OK' if ok else 'MISMATCH'}")
            bad += not ok
        except Exception as e:
            bad += 1
            print("handle API raised", e)
    # throughput probe
    for lv, total, chunk in ((1, 256 << 20, 1 << 20), (2, 128 << 20, 1 << 20), (3, 128 << 20, 1 << 20)):
        d = synth.tiled("mixed", total, tile=32 << 20)
        chunks = api.split_chunks(d, chunk)
        try:
            csc_b200.encode_chunks(chunks[:8], 1 << 30, lv)  # warm-up / allocations
            t = time.time()
            outs = csc_b200.encode_chunks(chunks, 1 << 30, lv)
            dt = time.time() - t
            st = csc_b200.get_stats()
            csz = sum(o.size for o in outs)
            print(f"throughput m{lv}: {total / dt / 1e6:.1f} MB/s e2e ({dt:.2f}s) ratio {csz / total:.4f} chunks {len(chunks)}")
            print("   stats", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()})
            # spot parity on 3 chunks
            for i in (0, len(chunks) // 2, len(chunks) - 1):
                r = R.encode(chunks[i], chunks[i].size, lv)
                ok = r.size == outs[i].size and (r == outs[i]).all()
                if not ok:
                    bad += 1
                    print(f"   chunk {i} DIFF@{first_diff(r, outs[i])}")
            t = time.time()
            back = csc_b200.decode_chunks(outs, [c.size for c in chunks])
            dt = time.time() - t
            okd = all(b.size == c.size and (b == c).all() for b, c in zip(back, chunks))
            print(f"   decode: {total / dt / 1e6:.1f} MB/s e2e roundtrip {'OK' if okd else 'MISMATCH'}")
            print("   dstats", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in csc_b200.get_stats().items() if v})
            bad += not okd
        except Exception as e:
            bad += 1
            print("throughput probe raised", e)
    print("TOTAL BAD", bad)


if __name__ == "__main__":
    main()
