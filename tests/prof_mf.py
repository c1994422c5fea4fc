"""ncu driver for the match-finder / analyzer kernels: one encode pass over N MiB (default 256) of text."""
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
