"""Mnemonic counts per kernel of the shipped library (cuobjdump -sass) -> profiles/*_sass_excerpts.txt.
usage: python tools/sass_counts.py [liblidarreg.so] > profiles/r2_sass_excerpts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lidarregistration_b200", "csrc", "liblidarreg.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "UBLKCP", "SYNCS", "ACQBULK", "FFMA2", "FFMA", "FMNMX3", "VHMNMX", "LEA.HI",
        "DFMA", "DMUL", "DADD", "MUFU", "HMNMX2", "LDG", "STG", "LDS", "STS", "ATOM", "ATOMS", "RED", "BAR.SYNC", "SHFL", "VOTE",
        "STL", "LDL"]
print("# cuobjdump -sass lidarregistration_b200/csrc/liblidarreg.so (sm_100a, the shipped build): mnemonic counts per kernel")
print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTCATOMSWS = tcgen05.alloc/dealloc, UBLKCP = cp.async.bulk,")
print("# SYNCS = mbarrier ops, ACQBULK = griddepcontrol.wait (programmatic dependent launch), FFMA2 = packed fma.rn.f32x2,")
print("# FMNMX3 / VHMNMX = 3-input min-max, DFMA/DMUL/DADD = fp64, STL/LDL = spills\n")
name, cnt, n = None, None, 0


def flush():
    if name:
        short = re.sub(r"^_Z\w*?\d+(k_[a-z0-9_]+)", r"\1", name)
        m = re.search(r"(k_[a-z0-9_]+)(I[A-Za-z0-9_]*E)?", name)
        label = m.group(1) if m else name[:40]
        t = re.search(r"IL[bi](\d)(?:EL[bi](\d))?E", name)
        if t:
            label += "<" + ",".join(x for x in t.groups() if x is not None) + ">"
        print("%-30s %5d instr  %s" % (label, n, "  ".join("%s %d" % (k, cnt[k]) for k in KEYS if cnt[k])))


for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name, cnt, n = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cnt is not None:
        n += 1
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                cnt[k] += 1
                break
flush()
