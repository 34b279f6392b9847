#!/usr/bin/env python
"""per kernel of libmodgpu.so: how many of the instructions that prove the Blackwell-native paths its SASS holds
(UBLKCP = cp.async.bulk / TMA, SYNCS = mbarrier, ATOMS / ATOMG = shared / global atomics, LDG/STG..., no HMMA / no library
kernels).  python tools/sass_excerpt.py > profiles/sass_r02.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "modimizer_b200/libmodgpu.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UBLKCP\S*|SYNCS\S*|ATOMS\S*|ATOMG\S*|RED\S*|UTMALDG\S*|HMMA\S*|UTC\w*MMA\S*|LDGSTS\S*|MATCH\S*|VOTE\S*|FLO\S*|BREV\S*|IMAD\.WIDE\S*|SHF\S*)")
kern, counts, total = None, collections.OrderedDict(), {}
arch = set()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        counts[kern] = collections.Counter(); total[kern] = 0
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    if kern and re.search(r"/\*[0-9a-f]{4}\*/", line):
        total[kern] += 1
        m = pat.search(line)
        if m:
            op = m.group(1).rstrip(";")
            op = ".".join(op.split(".")[:3]) if op.startswith(("ATOM", "SYNCS", "UBLKCP", "RED")) else op.split(".")[0]
            counts[kern][op] += 1
print("# %s: cubins for %s; %d kernels; instruction counts per kernel (static SASS)" % (lib, ", ".join(sorted(arch)), len(counts)))
keys = ["UBLKCP", "SYNCS", "ATOMS", "ATOMG", "RED"]
for k, c in counts.items():
    sel = {op: n for op, n in c.items() if op.startswith(tuple(keys))}
    short = k if len(k) < 110 else k[:107] + "..."
    print("%-112s %6d instr  %s" % (short, total[k], "  ".join("%s x%d" % (op, n) for op, n in sorted(sel.items()))))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("# totals:", "  ".join("%s x%d" % (op, n) for op, n in sorted(tot.items()) if op.startswith(tuple(keys + ["HMMA", "UTC", "UTMALDG"]))))
print("# tensor-core instructions (HMMA / UTC*MMA): %d - none by design: nothing on this path is a dense contraction" % sum(n for op, n in tot.items() if op.startswith(("HMMA", "UTC"))))
