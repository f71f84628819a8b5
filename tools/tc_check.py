import sys, ctypes; sys.path.insert(0, "/root/repo")
import torch, numpy as np
from bayesnn_fpga_b200 import _lib
from tests.test_gpu_ops import _conv_case, TC_SHAPES
_lib.build(); lib = _lib.load()
which = sys.argv[1:] 
for i, shape in enumerate(TC_SHAPES):
    if which and str(i) not in which: continue
    for dt in ("fp16",):
        got, want = _conv_case(lib, "tc", dt, *shape)
        d = (got.double() - want).abs()
        print(i, shape, dt, "max_err %.3e  scale %.2f  nan %d  frac_bad %.4f" % (
            np.nanmax(d.numpy()), want.abs().max().item(), int(torch.isnan(got).sum()),
            float((d > 1e-2 * max(1, want.abs().max().item())).float().mean())), flush=True)
        if np.nanmax(d.numpy()) > 1e-2 * max(1, want.abs().max().item()) or torch.isnan(got).any():
            bad = ((d > 1e-2) | torch.isnan(got)).nonzero()
            print("   first bad idx (n,c,h,w):", bad[:6].tolist(), " n_bad", len(bad))
            print("   got", got[tuple(bad[0])].item(), "want", want[tuple(bad[0])].item())
