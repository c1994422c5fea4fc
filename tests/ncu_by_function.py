"""Attribute the per-instruction counters of an ncu source-page CSV (SASS view) of k_encode to the device
functions inside the kernel, using nvdisasm's labels of the same build.
usage: python tests/ncu_by_function.py <ncu_source.csv> <libcsc_b200.so> [bytes_processed]"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
import os

src_csv, so = sys.argv[1], sys.argv[2]
nbytes = float(sys.argv[3]) if len(sys.argv) > 3 else 0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# instruction offset -> function label, inside the k_encode section
func_at = {}
cur, inside = None, False
for line in dis:
    m = re.match(r"^(\S+):\s*$", line)
    if m:
        lab = m.group(1)
        if lab.startswith(".L") or lab.startswith(".text."):
            continue
        if lab == "_ZN4cscb8k_encodeEPKNS_9EncStreamEjPjPKj":
            inside, cur = True, "k_encode(body: run/compress_segment inlined parts)"
        elif lab.startswith("$_ZN4cscb8k_encodeEPKNS_9EncStreamEjPjPKj$"):
            cur = lab.split("$")[2]
        elif not lab.startswith("$"):
            inside = False
        continue
    if inside:
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(\S.*?);", line)
        if m:
            func_at[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
inst, samp = collections.Counter(), collections.Counter()
for r in rows[2:]:
    try:
        a, n, s = int(r[ia], 16), int(r[ii]), int(r[isamp])
    except (ValueError, IndexError):
        continue
    if base is None:
        base = a
    f = func_at.get(a - base, "?")
    inst[f] += n
    samp[f] += s
ti, ts = sum(inst.values()), sum(samp.values())
print(f"total warp instructions {ti}" + (f" = {ti / nbytes:.1f} per input byte" if nbytes else ""))
try:
    dem = subprocess.run(["c++filt"] + list(inst.keys()), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(inst.keys(), dem))
except Exception:
    names = {}
for f, n in inst.most_common():
    print(f"{n / ti * 100:6.2f}% inst {samp[f] / max(ts, 1) * 100:6.2f}% samples  {names.get(f, f)[:110]}")
# optional: per-instruction listing of one function (4th argument = substring of its name)
if len(sys.argv) > 4:
    want = sys.argv[4]
    base = None
    for r in rows[2:]:
        try:
            a, n, s = int(r[ia], 16), int(r[ii]), int(r[isamp])
        except (ValueError, IndexError):
            continue
        if base is None:
            base = a
        f = func_at.get(a - base, "?")
        if want in f:
            print(f"{a - base:6x} {n / ti * 100:6.3f}% {s:7d}  {r[hdr.index('Source')].strip()[:100]}")
