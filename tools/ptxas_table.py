"""Turns paa_b200/libpaa_b200.so.ptxas.log (nvcc -Xptxas=-v, written by paa_b200/build.py) into the resource table
committed under profiles/: registers, stack, spills and static shared memory of every kernel.
    python tools/ptxas_table.py > profiles/r2_ptxas_resources.md"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
log = open(os.path.join(ROOT, "paa_b200", "libpaa_b200.so.ptxas.log")).read().split("\n")
rows, name = [], None
stack = spill_s = spill_l = 0
for line in log:
    m = re.search(r"Compiling entry function '([^']+)'", line)
    if m:
        name = m.group(1)
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        stack, spill_s, spill_l = (int(x) for x in m.groups())
        continue
    m = re.search(r"Used (\d+) registers(.*)", line)
    if m and name:
        smem = re.search(r"(\d+) bytes smem", m.group(2))
        rows.append((name, int(m.group(1)), stack, spill_s, spill_l, int(smem.group(1)) if smem else 0))
        name = None
names = subprocess.run(["c++filt"] + [r[0] for r in rows], stdout=subprocess.PIPE, text=True).stdout.strip().split("\n")
print("# ptxas -v resource usage of libpaa_b200.so (sm_100a), end of round 2 (`python tools/ptxas_table.py`)\n")
print("| kernel | registers | stack | spill stores | spill loads | static smem |\n|---|---|---|---|---|---|")
for (_, regs, st, ss, sl, smem), nm in zip(rows, names):
    nm = re.sub(r"\(.*$", "", nm).replace("void ", "").replace("paa::", "")
    print("| `%s` | %d | %d | %d | %d | %d |" % (nm, regs, st, ss, sl, smem))
