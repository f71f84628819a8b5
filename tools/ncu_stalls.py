#!/usr/bin/env python
"""Top stall sites of one profiled launch:  python tools/ncu_stalls.py REP.ncu-rep LAUNCH_INDEX [N] [sass|cuda]
(reads `ncu -i REP --page source --csv`; needs -lineinfo / --import-source on at capture time; `cuda` aggregates the
samples per CUDA-C source line)."""
import csv
import io
import subprocess
import sys

rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
view = sys.argv[4] if len(sys.argv) > 4 else "sass"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", view, "--launch-skip", str(idx),
                      "--launch-count", "1"],
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

# hot spots with their neighbourhood in program order (which wait loop a hot branch belongs to)
if view == "sass" and len(sys.argv) > 5 and sys.argv[5] == "ctx":
    print("\n--- hot SASS lines (>= 0.5 %% of samples) with the three lines in front of them ---")
    for i, r in enumerate(data):
        n = int(r[col["# Samples"]])
        if n * 200 >= tot:
            for j in range(max(0, i - 3), i + 1):
                q = data[j]
                print("%s %6d  %s" % (">>" if j == i else "  ", int(q[col["# Samples"]]), q[col["Source"]].strip()[:100]))
            print()

if view == "sass" and len(sys.argv) > 6:
    # program-order dump of a window around every line matching the pattern (e.g. UTCHMMA)
    pat = sys.argv[6]
    hits = [i for i, r in enumerate(data) if pat in r[col["Source"]]]
    if hits:
        lo, hi = max(0, hits[0] - 60), min(len(data), hits[-1] + 40)
        print("\n--- program order, lines %d..%d around %r ---" % (lo, hi, pat))
        for r in data[lo:hi]:
            why = sorted(((int(r[col[h]]), h) for h in stall_cols), reverse=True)[:1]
            print("%6d  %-90s %s" % (int(r[col["# Samples"]]), r[col["Source"]].strip()[:90],
                                     why[0][1] if why and why[0][0] else ""))
