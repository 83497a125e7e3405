#!/usr/bin/env python
"""SASS evidence for the built library: per kernel, instruction count, registers / shared memory (cuobjdump -res-usage) and
the mnemonics that prove the TMA / bulk-copy / mbarrier / 64-bit shared atomic claims of DESIGN.md (and that no tensor-core
instruction is present, as north_star states).  usage: tools/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cython3dmodelrenderer_b200", "csrc", "libcrender_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
WANT = ["UTMASTG", "UBLKCP", "SYNCS", "ATOMS.CAS.64", "ATOMS", "ATOMG", "RED.", "LDG.E.128", "LDS.128", "STS.128", "STG.E.128", "MUFU.RCP",
        "FFMA", "FMUL", "FADD", "BAR.SYNC", "S2R", "CCTL", "HMMA", "UTCMMA", "UTCHMMA"]
usage = {}
name = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        name = m.group(1); continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
    if m and name:
        usage[name] = tuple(int(x) for x in m.groups()); name = None
kern = collections.OrderedDict(); cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); kern[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1); kern[cur]["_n"] += 1
        for w in WANT:
            if op.startswith(w) or (w in ("ATOMS.CAS.64",) and w in op):
                kern[cur][w] += 1
def short(n):
    d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    m = re.search(r"(\w+)\(", d.replace("(anonymous namespace)::", ""))
    return m.group(1) if m else n
print(f"# {os.path.relpath(lib, ROOT)}: cuobjdump -sass / -res-usage, sm_100a")
arch = set(re.findall(r"arch = (sm_\w+)", sass)); print("# arch:", ", ".join(sorted(arch)))
tot = collections.Counter()
for k, c in kern.items():
    r = usage.get(k, (0, 0, 0))
    marks = "  ".join(f"{w}={c[w]}" for w in WANT if c[w])
    print(f"{short(k):26s} inst={c['_n']:5d} regs={r[0]:3d} stack={r[1]:3d} smem={r[2]:6d}  {marks}")
    tot.update(c)
print("# whole library: " + "  ".join(f"{w}={tot[w]}" for w in WANT))
print("# tensor-core instructions (HMMA / UTCMMA): %d -- none, by design (no contraction on this path)" % (tot["HMMA"] + tot["UTCMMA"] + tot["UTCHMMA"]))
