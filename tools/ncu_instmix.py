#!/usr/bin/env python
"""Instruction mix of one launch from `ncu -i REP --page source --csv --print-source sass` output:
python tools/ncu_instmix.py FILE.csv [tiles_per_sm]  -> executed warp instructions by opcode (and per tile)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
tiles = float(sys.argv[2]) if len(sys.argv) > 2 else None
head = rows[1]
col = {h: i for i, h in enumerate(head)}
data = [r for r in rows[2:] if len(r) == len(head) and r[col["Instructions Executed"]].isdigit()]
tot = sum(int(r[col["Instructions Executed"]]) for r in data)
print("warp instructions executed: %d (%.0f per SM%s)" % (tot, tot / 148, ", %.0f per tile" % (tot / 148 / tiles) if tiles else ""))
byop = collections.Counter()
for r in data:
    src = r[col["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    byop[m.group(2).split(".")[0] if m else src] += int(r[col["Instructions Executed"]])
for op, n in byop.most_common(28):
    print("%-12s %12d %5.1f%%%s" % (op, n, 100 * n / tot, "  %7.0f / tile" % (n / 148 / tiles) if tiles else ""))
