"""-m5 on a far-match workload (BASELINE config 4's recipe, scaled to what one serial stream allows): one 48 MiB stream of the
redundant corpus whose copies lie 20 MiB apart, coded by the host simulation of the bounded search (the GPU equals the host
simulation: test_m5_bounded_search_is_ratio_gated) and by libcsc -m5.  Takes ~15 minutes of CPU; result of 2026-10-17:
libcsc 11 328 719 bytes, bounded search 11 138 319 bytes (-1.68 %), decoded bit-exact by the reference decoder.
usage: python tests/m5_far_match_check.py [total MiB] [base MiB]"""
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from csc_b200 import synth
from oracle import refbind
from tests.hostsim import hsbind

total = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 48 << 20
base = int(sys.argv[2]) << 20 if len(sys.argv) > 2 else 20 << 20
d = synth.redundant(total, base=base, seed=11)
t = time.time(); r = refbind.encode(d, d.size, 5); t1 = time.time() - t
t = time.time(); h = hsbind.encode(d, d.size, 5); t2 = time.time() - t
back = refbind.decode(h, d.size)
print(f"libcsc -m5 {r.size} bytes ({t1:.0f} s), bounded search {h.size} bytes ({t2:.0f} s), delta {100.0 * (h.size - r.size) / r.size:+.3f} %, "
      f"reference decoder reads it: {back.size == d.size and bool((back == d).all())}")
