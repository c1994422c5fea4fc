"""Tool-level comparison on the box: tools/csarc_gpu.cpp (batched calls on the GPU) against the stock csarc with -t8 (its maximum)
on the same tree in /dev/shm.  usage: python tests/archiver_timing.py [MiB] [level]"""
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from csc_b200 import synth

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
level = sys.argv[2] if len(sys.argv) > 2 else "1"
base = "/dev/shm/csarc_timing"
shutil.rmtree(base, ignore_errors=True)
os.makedirs(base + "/src")
n = 0
for i in range(mib // 8):   # 8 MiB files: text, machine code, mixed
    kind = (synth.text, synth.exe, synth.mixed)[i % 3]
    kind(8 << 20, 100 + i).tofile(f"{base}/src/f{i:03d}.{('txt', 'exe', 'bin')[i % 3]}")
    n += 8 << 20
ref = os.path.join(ROOT, "oracle", "_ref")


def run(*a):
    t = time.time()
    r = subprocess.run(list(a), cwd=base, capture_output=True, text=True)
    assert r.returncode == 0, (a, r.stdout[-300:], r.stderr[-300:])
    return time.time() - t


run(f"{ref}/csarc_gpu", "a", f"-m{level}", "-c368k", "-f", "warm.csa", "src/f000.txt")   # CUDA context, buffers
tg = run(f"{ref}/csarc_gpu", "a", f"-m{level}", "-c368k", "-f", "gpu.csa", "src")
ts = run(f"{ref}/csarc_ref", "a", "-r", f"-m{level}", "-t8", "-f", "stock.csa", "src")
tx = run(f"{ref}/csarc_gpu", "t", "gpu.csa")
txs = run(f"{ref}/csarc_ref", "t", "-t8", "stock.csa")
run(f"{ref}/csarc_ref", "t", "-t8", "gpu.csa")
print(f"{n >> 20} MiB, -m{level}, {os.cpu_count()} host threads: csarc_gpu a {n / tg / 1e6:.0f} MB/s ({os.path.getsize(base + '/gpu.csa')} bytes), "
      f"stock csarc -t8 a {n / ts / 1e6:.0f} MB/s ({os.path.getsize(base + '/stock.csa')} bytes); "
      f"test: csarc_gpu {n / tx / 1e6:.0f} MB/s, stock -t8 {n / txs / 1e6:.0f} MB/s; stock csarc verifies the GPU archive")
shutil.rmtree(base, ignore_errors=True)
