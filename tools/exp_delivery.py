"""Operand-delivery experiment (measurement only): the same launches with the activation and / or weight TMA loads
switched off (BNN_TC_EXP bit 0 / bit 1).  If a launch speeds up to the MMA floor when its operands stop arriving, it
is paced by operand delivery into the SM, not by MMA issue or by its epilogue."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.conv_bench import bench
shapes = {"layer2 3x3 s1 128->128 (swapped)": (8192, 16, 16, 128, 128, 3, 1),
          "layer2.0 3x3 s2 64->128 (swapped)": (8192, 32, 32, 64, 128, 3, 2),
          "layer3 3x3 s1 256->256 (cta_group::2)": (8192, 8, 8, 256, 256, 3, 1),
          "layer3 3x3 s1 256->256 (single CTA)": (8192, 8, 8, 256, 256, 3, 1)}
for name, sh in shapes.items():
    if "single" in name:
        os.environ["BNN_TC_NOMC"] = "1"
    for exp in ("0", "1", "2", "3"):
        os.environ["BNN_TC_EXP"] = exp
        ms, tf = bench(*sh, iters=20)
        print("%-40s EXP=%s (%s) %.3f ms %7.1f TFLOP/s" % (name, exp, {"0": "all loads", "1": "no activation loads",
              "2": "no weight loads", "3": "no loads"}[exp], ms, tf), flush=True)
    os.environ.pop("BNN_TC_NOMC", None)
os.environ.pop("BNN_TC_EXP", None)
