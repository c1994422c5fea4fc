"""Per-kernel summary of an `ncu --page raw --csv` export: time, DRAM bytes and throughput, issue / occupancy, top stalls.
usage: python tests/ncu_kernel_table.py <raw.csv> [peak_GBps]"""
import csv
import sys

import numpy as np

rows = list(csv.reader(open(sys.argv[1])))
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6552.0
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "second": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}


def val(r, name):
    i = ix.get(name)
    if i is None or r[i] == "":
        return 0.0
    return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)


stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
agg = {}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    agg.setdefault(name, []).append(r)
print(f"{'kernel':24s} {'n':>3s} {'ms':>9s} {'rd GB':>8s} {'wr GB':>8s} {'DRAM GB/s':>10s} {'% peak':>7s} {'issue%':>7s} {'occ%':>6s} {'regs':>5s}  top stalls (cycles per issued instruction)")
for name, rs in agg.items():
    t = sum(val(r, "gpu__time_duration.sum") for r in rs)
    rd = sum(val(r, "dram__bytes_read.sum") for r in rs)
    wr = sum(val(r, "dram__bytes_write.sum") for r in rs)
    w = np.array([val(r, "gpu__time_duration.sum") for r in rs])
    w = w / w.sum() if w.sum() else w
    issue = float(np.dot([val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") for r in rs], w))
    occ = float(np.dot([val(r, "sm__warps_active.avg.pct_of_peak_sustained_active") for r in rs], w))
    st = {s[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: float(np.dot([val(r, s) for r in rs], w)) for s in stalls}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
    gbs = (rd + wr) / t / 1e9 if t else 0
    print(f"{name[:24]:24s} {len(rs):3d} {t * 1e3:9.3f} {rd / 1e9:8.2f} {wr / 1e9:8.2f} {gbs:10.1f} {gbs / peak * 100:7.1f} {issue:7.1f} {occ:6.1f} {val(rs[0], 'launch__registers_per_thread'):5.0f}  "
          + ", ".join(f"{k} {v:.1f}" for k, v in top))
