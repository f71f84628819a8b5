import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from bayesnn_fpga_b200.utils import Masksembles2D
np.random.seed(0)
m2 = Masksembles2D(64, 4, 2.0).cuda().eval()
x = torch.randn(8, 64, 3, 3, device="cuda")
for call in range(3):
    got = m2(x)
    want = x * m2.masks[call % 4].view(1, -1, 1, 1)
    bad = (got != want)
    print(call, "mismatches", int(bad.sum()), "of", bad.numel())
    if bad.any():
        idx = bad.nonzero()[:5]
        print(idx.tolist(), got[bad][:5].tolist(), want[bad][:5].tolist())
        # which row would match?
        for r in range(4):
            print("row", r, int((got != x * m2.masks[r].view(1,-1,1,1)).sum()))
