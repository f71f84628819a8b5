"""Back-off in the mbarrier waits (BNN_TC_WAIT_NS = "producer,epilogue,accumulator,operands" in ns; 0 = tight try_wait
loop).  Interleaved, minimum of 6 rounds.  Cases: layer2.0.1.conv2 (vertical halo + residual + fused element dropout),
layer2.0.1.conv1 (vertical halo, plain), sibling group 1 (sibling pair), a 256-channel 8x8 conv (cta_group::2)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesnn_fpga_b200 import _lib
from tests.gpu_util import drop_desc, stream
lib = _lib.load()
N = 8192


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


def conv_case(H, Cin, Cout, stride, res, drop):
    x = torch.randn(N, H, H, Cin, device="cuda", dtype=torch.float16)
    w = (torch.randn(Cout, 3, 3, Cin, device="cuda") / (Cin * 9) ** 0.5).half()
    b = torch.randn(Cout, device="cuda")
    OH = H // stride
    y = torch.empty(N, OH, OH, Cout, device="cuda", dtype=torch.float16)
    r = torch.randn_like(y) if res else None
    dd = drop_desc(1, 0.25, 0x77, 3, 0, 256) if drop else drop_desc(batch=N)
    keep = (x, w, b, y, r, dd)
    return lambda: lib.bnn_conv2d_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), ctypes.c_void_p(r.data_ptr() if res else 0), y.data_ptr(),
                                     1, N, H, H, Cin, Cout, 3, stride, 1, ctypes.byref(dd), stream()), keep


def sibling_case():
    x2 = torch.randn(N, 32, 32, 64, device="cuda", dtype=torch.float16)
    w2 = (torch.randn(256, 3, 3, 64, device="cuda") / 24).half()
    b2 = torch.randn(256, device="cuda")
    outs = [torch.empty(N, 16, 16, 128, device="cuda", dtype=torch.float16) for _ in range(2)]
    ys = (ctypes.c_void_p * 2)(*[o.data_ptr() for o in outs])
    keep = (x2, w2, b2, outs, ys)
    return lambda: lib.bnn_conv2d_tc_grouped(x2.data_ptr(), w2.data_ptr(), b2.data_ptr(), ys, 2, 3, 0, 1, N, 32, 32, 64, 128, 3, 2,
                                             stream()), keep


cases = {"L10 vh+res+dropout": conv_case(16, 128, 128, 1, True, True), "L9 vh plain": conv_case(16, 128, 128, 1, False, False),
         "sibling pair K=576": sibling_case(), "8x8 256ch cg2": conv_case(8, 256, 256, 1, False, False),
         "8x8 256ch cg2 +res+drop": conv_case(8, 256, 256, 1, True, True), "4x4 512ch pm": conv_case(4, 512, 512, 1, False, False)}
variants = sys.argv[1:] or ["0,0,0,0", "64,64,32,0", "64,0,0,0", "0,64,0,0", "0,0,32,0", "128,128,64,0", "64,64,32,20", "32,32,32,0"]
best = {}
for rnd in range(6):
    for cname, (call, _) in cases.items():
        for v in variants:
            os.environ["BNN_TC_WAIT_NS"] = v
            best[(cname, v)] = min(best.get((cname, v), 1e9), timed(call))
for cname in cases:
    print(cname)
    for v in variants:
        print("   wait_ns %-14s %.4f ms" % (v, best[(cname, v)]), flush=True)
