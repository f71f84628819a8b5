"""Timing of the 256-column kernel with a fused stochastic site (layer3.0.1.conv2 of C2) vs the plain convolution.
Interleaved, minimum of 6 rounds."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesnn_fpga_b200 import _lib
from tests.gpu_util import drop_desc, stream
lib = _lib.load()
N = 8192


def case(H, C, res, drop):
    x = torch.randn(N, H, H, C, device="cuda", dtype=torch.float16)
    w = (torch.randn(C, 3, 3, C, device="cuda") / (C * 9) ** 0.5).half()
    b = torch.randn(C, device="cuda")
    y = torch.empty(N, H, H, C, device="cuda", dtype=torch.float16)
    r = torch.randn_like(y) if res else None
    dd = drop_desc(1, 0.25, 0x77, 3, 0, 256) if drop else drop_desc(batch=N)
    keep = (x, w, b, y, r, dd)
    return lambda: lib.bnn_conv2d_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), ctypes.c_void_p(r.data_ptr() if res else 0), y.data_ptr(),
                                     1, N, H, H, C, C, 3, 1, 1, ctypes.byref(dd), stream()), keep


def timed(call, iters=20):
    for _ in range(3):
        assert call() == 0, lib.bnn_last_error()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


cases = {"8x8 256ch plain": case(8, 256, False, False), "8x8 256ch +res": case(8, 256, True, False),
         "8x8 256ch +res +dropout": case(8, 256, True, True), "4x4 512ch +res +dropout": case(4, 512, True, True),
         "4x4 512ch plain": case(4, 512, False, False)}
best = {}
for rnd in range(6):
    for name, (call, _) in cases.items():
        for epi in ("16", "8"):
            os.environ["BNN_TC_EPI256"] = epi
            best[(name, epi)] = min(best.get((name, epi), 1e9), timed(call))
for name in cases:
    print("%-28s epilogue warps 16: %.4f ms   8: %.4f ms" % (name, best[(name, "16")], best[(name, "8")]), flush=True)
