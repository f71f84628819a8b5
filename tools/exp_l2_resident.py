"""Is the operand-swapped kernel paced by DRAM-miss latency at tile switches or by the activation-tile delivery rate?
Same layer with inputs that fit the 126 MB L2 (N <= 1024 images of 16x16x128 fp16 = 64 KB each) vs the full 8192."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.conv_bench import bench
for name, sh in (("layer2 3x3 s1 128->128", (16, 16, 128, 128, 3, 1)), ("layer2.0 3x3 s2 64->128", (32, 32, 64, 128, 3, 2)),
                 ("layer3 3x3 s1 256->256", (8, 8, 256, 256, 3, 1))):
    for N in (592, 1184, 2368, 8288):          # multiples of 148 tiles (one 256-pixel tile per image at 16x16 outputs)
        for exp in ("0", "1"):
            os.environ["BNN_TC_EXP"] = exp
            ms, tf = bench(N, *sh, iters=20)
            print("%-26s N=%5d EXP=%s %.4f ms %7.1f TFLOP/s" % (name, N, exp, ms, tf), flush=True)
os.environ.pop("BNN_TC_EXP", None)
