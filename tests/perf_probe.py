"""Throughput probe for a gpurun call: python tests/perf_probe.py kind:level:totalMiB:chunkKiB ..."""
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np

import csc_b200
from csc_b200 import api, synth

for spec in sys.argv[1:]:
    kind, level, total, chunk = spec.split(":")
    level, total, chunk = int(level), int(total) << 20, int(chunk) << 10
    data = synth.tiled(kind, total, tile=32 << 20)
    chunks = api.split_chunks(data, chunk)
    csc_b200.encode_chunks(chunks[:4], 1 << 30, level)
    t = time.time()
    outs = csc_b200.encode_chunks(chunks, 1 << 30, level)
    dt = time.time() - t
    st = csc_b200.get_stats()
    comp = sum(o.size for o in outs)
    enc = {k[3:]: round(v, 1) for k, v in st.items() if k.startswith("ms_") and v > 0.5}
    enc["waves"] = st.get("waves")
    t = time.time()
    back = csc_b200.decode_chunks(outs, [c.size for c in chunks])
    dd = time.time() - t
    ds = csc_b200.get_stats()
    ok = all(b.size == c.size and (b == c).all() for b, c in zip(back, chunks))
    print(f"{spec}: enc e2e {total / dt / 1e6:.0f} MB/s gpu {total / st['ms_total_gpu'] / 1e3:.0f} MB/s ratio {comp / total:.4f} {enc} | "
          f"dec e2e {total / dd / 1e6:.0f} MB/s k_decode {ds['ms_decode']:.0f} ms unfilter {ds['ms_unfilter']:.0f} ms roundtrip {'OK' if ok else 'BAD'}",
          flush=True)
