"""Attribute the per-instruction counters of an ncu source-page CSV (SASS view) of one kernel to SOURCE LINES,
using nvdisasm -g line info of the same build.
usage: python tests/ncu_by_line.py <ncu_source.csv> <libcsc_b200.so> <kernel mangled name> [top N] [bytes]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

src_csv, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
nbytes = float(sys.argv[5]) if len(sys.argv) > 5 else 0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-c", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
line_at, func_at = {}, {}
cur_line, cur_fn, inside = "?", "?", False
for l in dis:
    m = re.match(r"^(\S+):\s*$", l)
    if m:
        lab = m.group(1)
        if lab == kern:
            inside, cur_fn = True, "(kernel body)"
        elif lab.startswith("$" + kern + "$"):
            cur_fn = lab.split("$")[2]
        elif not lab.startswith(".") and not lab.startswith("$"):
            inside = False
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = os.path.basename(m.group(1)) + ":" + m.group(2)
        continue
    if inside:
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
        if m:
            a = int(m.group(1), 16)
            line_at[a] = cur_line
            func_at[a] = cur_fn
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
inst, samp, fn_of = collections.Counter(), collections.Counter(), {}
for r in rows[2:]:
    try:
        a, n, s = int(r[ia], 16), int(r[ii]), int(r[isamp])
    except (ValueError, IndexError):
        continue
    if base is None:
        base = a
    k = line_at.get(a - base, "?")
    inst[k] += n
    samp[k] += s
    fn_of[k] = func_at.get(a - base, "?")
ti, ts = sum(inst.values()), sum(samp.values())
print(f"total warp instructions {ti}" + (f" = {ti / nbytes:.1f} per input byte" if nbytes else ""))
names = {}
try:
    fl = sorted(set(fn_of.values()))
    dem = subprocess.run(["c++filt"] + fl, capture_output=True, text=True).stdout.splitlines()
    names = {a: b.split("(")[0].replace("cscb::Encoder::", "").replace("cscb::", "") for a, b in zip(fl, dem)}
except Exception:
    pass
for k, n in sorted(inst.items(), key=lambda kv: -samp[kv[0]])[:topn]:
    print(f"{samp[k] / max(ts, 1) * 100:6.2f}% samples {n / ti * 100:6.2f}% inst  {k:28s} in {names.get(fn_of[k], fn_of[k])}")
