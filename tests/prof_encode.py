"""Minimal driver for ncu captures: one device-resident encode pass over N MiB of the bench text.
usage: python tests/prof_encode.py [total_MiB] [level] [chunk_KiB] [decode]"""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import csc_b200
from csc_b200 import api, synth

total = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 1 << 30
level = int(sys.argv[2]) if len(sys.argv) > 2 else 2
chunk = int(sys.argv[3]) << 10 if len(sys.argv) > 3 else 1 << 20
kind = sys.argv[4] if len(sys.argv) > 4 else "text"
tile = 32 << 20
data = np.empty(total, np.uint8)
for i, off in enumerate(range(0, total, tile)):
    n = min(tile, total - off)
    data[off:off + n] = {"text": synth.text, "mixed": synth.mixed, "exe": synth.exe}[kind](n, 12345 + i)
d = torch.from_numpy(data).cuda()
n = (total + chunk - 1) // chunk
ptrs = [d.data_ptr() + i * chunk for i in range(n)]
sizes = [min(chunk, total - i * chunk) for i in range(n)]
out = api.encode_chunks_device(ptrs, sizes, 64 << 20, level)
st = csc_b200.get_stats()
print("ratio", sum(out) / total, {k: round(v, 2) for k, v in st.items() if k.startswith("ms_") and v})
