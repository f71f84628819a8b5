"""Resident weights in the sibling-pair kernel (BNN_TC_NO_RW=1: weights re-fetched per k-block). Interleaved, min of 8."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesnn_fpga_b200 import _lib
from tests.gpu_util import stream
lib = _lib.load()
N = 8192
x2 = torch.randn(N, 32, 32, 64, device="cuda", dtype=torch.float16)
w2 = (torch.randn(256, 3, 3, 64, device="cuda") / 24).half()
b2 = torch.randn(256, device="cuda")
outs = [torch.empty(N, 16, 16, 128, device="cuda", dtype=torch.float16) for _ in range(2)]
ys = (ctypes.c_void_p * 2)(*[o.data_ptr() for o in outs])
call = lambda: lib.bnn_conv2d_tc_grouped(x2.data_ptr(), w2.data_ptr(), b2.data_ptr(), ys, 2, 3, 0, 1, N, 32, 32, 64, 128, 3, 2, stream())


def timed(iters=20):
    for _ in range(3):
        assert call() == 0, lib.bnn_last_error()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


variants = [("resident weights", {}), ("resident + L2 prefetch 1 ahead", {"BNN_TC_L2PF": "1"}), ("resident + L2 prefetch 2 ahead", {"BNN_TC_L2PF": "2"}),
            ("per k-block + L2 prefetch 1 ahead", {"BNN_TC_NO_RW": "1", "BNN_TC_L2PF": "1"}), ("weights per k-block", {"BNN_TC_NO_RW": "1"}), ("one CTA per (tile, group)", {"BNN_TC_NO_SCG2": "1"})]
best, ref = {}, {}
for rnd in range(8):
    for name, env in variants:
        os.environ.update(env)
        best[name] = min(best.get(name, 1e9), timed())
        ref[name] = [o.clone() for o in outs]
        for k in env:
            os.environ.pop(k)
fl = 2 * N * 256 * 256 * 64 * 9
for name, _ in variants:
    same = all(torch.equal(a, b) for a, b in zip(ref[name], ref["one CTA per (tile, group)"]))
    print("%-28s min of 8: %.4f ms %7.1f TFLOP/s   bit-identical to the single-CTA kernel: %s" % (name, best[name], fl / best[name] / 1e9, same), flush=True)
