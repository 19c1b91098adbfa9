"""SASS evidence for profiles/: mnemonic counts per kernel of libpaa_b200.so (cuobjdump -sass).
python tools/sass_summary.py > profiles/r2_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "paa_b200", "libpaa_b200.so")
WATCH = ["LDG", "STG", "LDS", "STS", "RED", "ATOM", "ATOMS", "REDUX", "SHFL", "VOTE", "MATCH", "MUFU", "FFMA", "DFMA", "DADD", "DMUL",
         "F2F", "UBLKPF", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "HMMA", "DMMA", "UTCHMMA", "LDTM", "BAR", "MEMBAR",
         "ACQBULK", "CCTL"]
out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
kern, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        kern[cur][m.group(1)] += 1
        kern[cur]["_total"] += 1


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip().split("(")[0].replace("paa::", "")
    except OSError:
        return n


print("# SASS summary of paa_b200/libpaa_b200.so (`cuobjdump -sass`), cubins for: %s\n" % ", ".join(arch))
tot = collections.Counter()
for c in kern.values():
    tot.update(c)
print("Whole library: %d kernels, %d instructions.  " % (len(kern), tot["_total"]) +
      ", ".join("%s %d" % (k, tot[k]) for k in WATCH if tot[k]) + ".")
print("No tcgen05 / TMA tensor instructions (UTC*MMA, UTMALDG, LDTM): nothing on the path is a contraction, and the "
      "streaming kernels reach their bandwidth with plain vector loads (DESIGN.md 4).  `UBLKPF` is "
      "`cp.async.bulk.prefetch.L2` (bulk_focal_kernel), `REDUX` the warp-wide integer reductions (IoU matching), "
      "`DFMA` the float64 EM chain.\n")
cols = ["_total"] + [k for k in WATCH if tot[k]]
print("| kernel | " + " | ".join("instr" if c == "_total" else c for c in cols) + " |")
print("|---|" + "---|" * len(cols))
for name, c in kern.items():
    print("| %s | " % demangle(name) + " | ".join(str(c[k]) if c[k] else "" for k in cols) + " |")
