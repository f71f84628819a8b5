#!/usr/bin/env python
"""SASS mnemonic counts per kernel of libbnn_b200.so -> profiles/rNN_sass_counts.txt
(python tools/sass_counts.py > profiles/r02_sass_counts.txt; cuobjdump only, no GPU)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "bayesnn_fpga_b200/csrc/libbnn_b200.so"
COLS = ["UTCHMMA", "UTCIMMA", "2CTA", "LDTM", "UTMALDG", "MULTICAST", "UTCBAR", "ELECT", "VIMNMX", "HMMA", "UTMASTG", "STTM"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
counts, order, cur, k = {}, [], None, 0
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = names[k] if k < len(names) and names[k] else m.group(1)
        k += 1
        cur = re.sub(r"\(CUtensorMap_st.*", "", cur)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    if cur is None or "/*" not in line:
        continue
    for c in COLS:
        if c == "2CTA":
            if re.search(r"UTC[HI]MMA\.2CTA", line):
                counts[cur][c] += 1
        elif c == "HMMA":
            if re.search(r"\bHMMA\.", line):
                counts[cur][c] += 1
        elif c == "VIMNMX":
            if "VIMNMX.U16x2" in line:
                counts[cur][c] += 1
        elif c in line:
            counts[cur][c] += 1
print("# SASS instruction counts per kernel (cuobjdump -sass %s, sm_100a; tools/sass_counts.py)" % lib)
print("# UTCHMMA = tcgen05.mma kind::f16, UTCIMMA = kind::i8, 2CTA = cta_group::2 MMAs, LDTM = tcgen05.ld, UTMALDG = TMA tensor")
print("# load, UTCBAR = tcgen05.commit, VIMNMX = the two-halfword threshold compares of the Philox keep bits, HMMA = legacy")
print("# mma.sync (wide exit heads on bf16 / F % 64 != 0), UTMASTG = TMA store (none: epilogues store from registers, DESIGN.md 9)")
print()
print("%-104s" % "kernel" + "".join("%10s" % c for c in COLS))
tot = collections.Counter()
for n in order:
    if not any(counts[n][c] for c in COLS):
        continue
    print("%-104s" % n[:104] + "".join("%10d" % counts[n][c] for c in COLS))
    tot.update(counts[n])
print("%-104s" % "TOTAL" + "".join("%10d" % tot[c] for c in COLS))
