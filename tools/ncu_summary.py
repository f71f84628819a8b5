#!/usr/bin/env python
"""Summarise an ncu report for profiles/:  python tools/ncu_summary.py REP.ncu-rep OUT.json [name ...]

Reads `ncu -i REP --page raw --csv` (no GPU needed) and keeps, per profiled launch: kernel, duration, tensor-pipe
activity, DRAM bytes read / written, L2 hit rate, SM clock, registers.  Optional names label the launches in order."""
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "us",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_hmma_pct",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__cycles_elapsed.avg.per_second": "sm_ghz",
    "launch__registers_per_thread": "regs",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
}


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def main():
    rep, out = sys.argv[1], sys.argv[2]
    names = sys.argv[3:]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    res = []
    for k, r in enumerate(data):
        d = {"launch": k, "kernel": r[col["Kernel Name"]][:60]}
        if k < len(names):
            d["layer"] = names[k]
        for m, short in KEEP.items():
            if m not in col:
                continue
            v, u = to_float(r[col[m]]), units[col[m]]
            if v is None:
                continue
            if short == "us":
                v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
            elif short.endswith("_MB"):
                v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
            elif short == "sm_ghz":
                v *= {"hz": 1e-9, "Khz": 1e-6, "Mhz": 1e-3, "Ghz": 1.0, "cycle/second": 1e-9, "cycle/nsecond": 1.0,
                      "cycle/usecond": 1e-3}.get(u, 1e-9)
            d[short] = round(v, 3)
        res.append(d)
    json.dump(res, open(out, "w"), indent=0)
    print("wrote %d launches to %s" % (len(res), out))


if __name__ == "__main__":
    main()
