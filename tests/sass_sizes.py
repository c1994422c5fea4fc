"""Instruction count and local-memory (LDL/STL) count of every out-of-line device function of one kernel in the built library.
usage: python tests/sass_sizes.py [libcsc_b200.so] [kernel mangled name]"""
import os, re, subprocess, sys, tempfile
so = os.path.abspath(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "csc_b200", "csrc", "libcsc_b200.so"))
kern = sys.argv[2] if len(sys.argv) > 2 else "_ZN4cscb8k_encodeEPKNS_9EncStreamEjPjPKj"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
lines = subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
inside, cur, sizes, ldl, order = False, None, {}, {}, []
for l in lines:
    m = re.match(r"^(\S+):\s*$", l)
    if m:
        lab = m.group(1)
        if lab.startswith(".L") or lab.startswith(".text."):
            continue
        if lab == kern:
            inside, cur = True, "(kernel body)"
        elif lab.startswith("$" + kern + "$"):
            cur = lab.split("$")[2]
        elif not lab.startswith("$"):
            inside = False
            continue
        if inside and cur not in sizes:
            order.append(cur); sizes[cur] = 0; ldl[cur] = 0
        continue
    if inside and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l):
        sizes[cur] += 1
        if re.search(r"\b(LDL|STL)", l):
            ldl[cur] += 1
names = subprocess.run(["c++filt"] + order, capture_output=True, text=True).stdout.splitlines()
print(f"{'instr':>6} {'LDL/STL':>7}  function   (total {sum(sizes.values())} instructions = {sum(sizes.values()) * 16 >> 10} KiB)")
for c, d in zip(order, names):
    print(f"{sizes[c]:6d} {ldl[c]:7d}  {d[:110]}")
