"""Vertical-halo form of the operand-swapped kernel: ring split and operand-delivery switches (measurement only).
Variants are interleaved over several rounds and the MINIMUM time per variant is reported (clock / power drift between
back-to-back measurements is larger than the differences of interest)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.conv_bench import bench
sh = (8192, 16, 16, 128, 128, 3, 1)
variants = [("per-tap kernel", {"BNN_TC_NO_VH": "1"}), ("per-tap, no act loads", {"BNN_TC_NO_VH": "1", "BNN_TC_EXP": "1"}),
            ("VH 3,7", {}), ("VH 3,5", {"BNN_TC_VH_SLOTS": "3,5"}), ("VH 4,4", {"BNN_TC_VH_SLOTS": "4,4"}),
            ("VH 2,9", {"BNN_TC_VH_SLOTS": "2,9"}), ("VH 2,6", {"BNN_TC_VH_SLOTS": "2,6"}),
            ("VH 3,7 no act loads", {"BNN_TC_EXP": "1"}), ("VH 3,7 no weight loads", {"BNN_TC_EXP": "2"}),
            ("VH 3,7 no loads", {"BNN_TC_EXP": "3"})]
best = {}
for rnd in range(6):
    for name, env in variants:
        os.environ.update(env)
        ms, tf = bench(*sh, iters=30)
        for k in env:
            os.environ.pop(k)
        best[name] = min(best.get(name, 1e9), ms)
fl = 2 * 8192 * 256 * 128 * 128 * 9
for name, _ in variants:
    print("%-26s min of 6: %.4f ms %7.1f TFLOP/s" % (name, best[name], fl / best[name] / 1e9), flush=True)
