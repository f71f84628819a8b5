#!/usr/bin/env python
"""Top stall sites of one profiled launch:  python tools/ncu_stalls.py REP.ncu-rep LAUNCH_INDEX [N]
(reads `ncu -i REP --page source --csv`; needs -lineinfo / --import-source on at capture time)."""
import csv
import io
import subprocess
import sys

rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
print(rows[0][1][:120])
head = rows[1]
col = {h: i for i, h in enumerate(head)}
stall_cols = [h for h in head if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) != len(head) or not r[col["# Samples"]].isdigit():
        break                      # the SASS view comes first; stop at the next view's header
    data.append(r)
tot = sum(int(r[col["# Samples"]]) for r in data)
agg = {h: sum(int(r[col[h]]) for r in data) for h in stall_cols}
print("samples", tot, {k: round(v / tot, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(data, key=lambda r: -int(r[col["# Samples"]]))[:top]:
    why = sorted(((int(r[col[h]]), h) for h in stall_cols), reverse=True)[:2]
    print("%6d %5.1f%%  %-70s %s" % (int(r[col["# Samples"]]), 100.0 * int(r[col["# Samples"]]) / tot, r[col["Source"]].strip()[:70],
                                     ", ".join("%s=%d" % (h, v) for v, h in why)))
