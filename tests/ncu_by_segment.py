"""Attribute an ncu source-page CSV (SASS view) to code segments delimited by CALL targets (= out-of-line device functions),
without needing the binary.  usage: python tests/ncu_by_segment.py <ncu_source.csv> [bytes] [dump_segment_index]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
nbytes = float(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = rows[1]
ia, isrc, ii, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ins = []
for r in rows[2:]:
    try:
        ins.append((int(r[ia], 16), r[isrc].strip(), int(r[ii]), int(r[isamp])))
    except (ValueError, IndexError):
        pass
base = ins[0][0]
targets = {base}
for a, s, n, sm in ins:
    m = re.search(r"CALL\.REL\.NOINC (0x[0-9a-f]+)", s)
    if m:
        targets.add(int(m.group(1), 16))
starts = sorted(targets)
import bisect
seg = collections.defaultdict(lambda: [0, 0, 0, 0])   # inst, samples, size, calls
calls = collections.Counter()
for a, s, n, sm in ins:
    k = starts[bisect.bisect_right(starts, a) - 1]
    seg[k][0] += n; seg[k][1] += sm; seg[k][2] += 1
    m = re.search(r"CALL\.REL\.NOINC (0x[0-9a-f]+)", s)
    if m:
        calls[int(m.group(1), 16)] += n
ti = sum(v[0] for v in seg.values()); ts = sum(v[1] for v in seg.values())
print(f"total warp instructions {ti}" + (f" = {ti / nbytes:.1f} per input byte" if nbytes else ""), "samples", ts)
for i, k in enumerate(starts):
    v = seg[k]
    if v[0] == 0: continue
    print(f"seg {i:3d} @{k - base:6x} size {v[2]:5d}  {v[0] / ti * 100:6.2f}% inst {v[1] / max(ts, 1) * 100:6.2f}% samples  calls {calls[k]:>11d}  inst/call {v[0] / max(calls[k], 1):8.1f}")
if len(sys.argv) > 3:
    want = starts[int(sys.argv[3])]
    nxt = starts[int(sys.argv[3]) + 1] if int(sys.argv[3]) + 1 < len(starts) else 1 << 62
    for a, s, n, sm in ins:
        if want <= a < nxt:
            print(f"{a - base:6x} {n:>11d} {sm:7d}  {s[:110]}")
