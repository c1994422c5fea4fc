"""ncu driver for the parallel kernels: one encode pass over N MiB (default 256) of text (match-finder / analyzer / sort kernels),
a second one over 64 MiB of the mixed corpus at -m3 (k_mf_tlist, filters on machine code), then Adler-32 and a decode.
usage: python tests/prof_mf.py [total MiB] [level]"""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import csc_b200
from csc_b200 import api, synth

total = (int(sys.argv[1]) if len(sys.argv) > 1 else 256) << 20
level = int(sys.argv[2]) if len(sys.argv) > 2 else 2
chunk = 512 << 10
data = synth.tiled("text", total, tile=32 << 20)
d = torch.from_numpy(data).cuda()
n = total // chunk
out = api.encode_chunks_device([d.data_ptr() + i * chunk for i in range(n)], [chunk] * n, 64 << 20, level)
print("ratio", sum(out) / total, {k: round(v, 2) for k, v in csc_b200.get_stats().items() if k.startswith("ms_") and v})

mixed = synth.tiled("mixed", 64 << 20, tile=32 << 20)
chunks = api.split_chunks(mixed, chunk)
outs, sums = api.encode_chunks_adler(chunks, 64 << 20, 3)
back, sums2 = api.decode_chunks_adler(outs, [c.size for c in chunks])
print("mixed -m3 ratio", sum(o.size for o in outs) / mixed.size, "adler ok", sums == sums2, "roundtrip", all((b == c).all() for b, c in zip(back, chunks)))
