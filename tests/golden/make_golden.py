"""Generates tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run in the development container:
    python tests/golden/make_golden.py
Inputs are the deterministic synth.edge_cases(); for each (case, level) we pin size + sha256 of
the reference stream, and a few tiny streams are stored verbatim."""
import hashlib
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from csc_b200 import synth  # noqa: E402
from oracle import refbind as R  # noqa: E402

ec = synth.edge_cases()
out = {"generator": "tests/golden/make_golden.py", "reference": "fusiyuan2010/CSC libcsc (zero-filling ISzAlloc)",
       "streams": [], "small_streams": []}
for name, d in ec.items():
    for level in (1, 2, 3, 4, 5):
        s = R.encode(d, max(d.size, 32768), level)
        out["streams"].append({"name": name, "level": level, "size": int(s.size),
                               "input_sha256": hashlib.sha256(d.tobytes()).hexdigest(),
                               "sha256": hashlib.sha256(s.tobytes()).hexdigest()})
for name in ("empty", "one", "tiny"):
    for level in (1, 3):
        d = ec[name]
        s = R.encode(d, 32768, level)
        out["small_streams"].append({"name": name, "level": level, "input_hex": d.tobytes().hex(), "stream_hex": s.tobytes().hex()})
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "golden.json"), "w"), indent=1)
print("wrote", len(out["streams"]), "stream hashes")
