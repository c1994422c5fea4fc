import numpy as np, sys
sys.path.insert(0, '.')
import csc_b200
from csc_b200 import api, synth
d = synth.mixed(64 << 20, 77)
chunks = api.split_chunks(d, 368 << 10)
one = csc_b200.encode_chunks(chunks, 64 << 20, 1)
csc_b200.set_devices([0, 1])
two = csc_b200.encode_chunks(chunks, 64 << 20, 1)
back = csc_b200.decode_chunks(two, [c.size for c in chunks])
print("fan-out over 2 GPUs: identical to one GPU:", all(a.size == b.size and (a == b).all() for a, b in zip(one, two)),
      "round trip:", all((b == c).all() for b, c in zip(back, chunks)), "devices", csc_b200.device_count())
