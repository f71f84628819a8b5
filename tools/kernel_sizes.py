#!/usr/bin/env python
"""Code size (bytes) of every conv_tc_kernel instantiation in libbnn_b200.so (the fully unrolled epilogues must stay well
inside the instruction cache: a 100 KB epilogue ran 3x slower)."""
import re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "bayesnn_fpga_b200/csrc/libbnn_b200.so"
out = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
rows = []
for line in out.splitlines():
    m = re.match(r"\s*[0-9a-f]+\s+[0-9a-f]+\s+([0-9a-f]+)\s.*\.text\.(\S+)", line)
    if m and "conv_tc_kernel" in m.group(2):
        name = re.sub(r"^_ZN3bnn2tc14conv_tc_kernelI", "", m.group(2))
        rows.append((int(m.group(1), 16), name[:60]))
for sz, n in sorted(rows):
    print("%7d  %s" % (sz, n))
