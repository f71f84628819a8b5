"""Masked sibling-pair kernel: where does the time go? (measurement switches BNN_TC_EXP: 4 = no RMW, 8 = no proxy fence)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesnn_fpga_b200 import _lib
from tests.gpu_util import drop_desc, stream
lib = _lib.load()
B, S, H, C, G = 256, 32, 32, 64, 2
x = torch.randn(B, H, H, C).relu().half().cuda()
dd = drop_desc(1, 0.5, 1, 2, 0, B)
xs = torch.empty_like(x)
bits = torch.zeros(S * B * H * H * C // 8, dtype=torch.uint8, device="cuda")
plane = torch.empty(S * B, H // 2, H // 2, C, dtype=torch.half, device="cuda")
copies = torch.empty(S * B, H, H, C, dtype=torch.half, device="cuda")
lib.bnn_dropout(x.data_ptr(), copies.data_ptr(), 1, H * H * C, C, S, 0, ctypes.byref(dd), stream())
lib.bnn_boundary_bits(x.data_ptr(), xs.data_ptr(), bits.data_ptr(), plane.data_ptr(), 1, B, H, H, C, S, ctypes.byref(dd), stream())
w = (torch.randn(G * 128, 3, 3, C) / 24).half().cuda()
b = torch.randn(G * 128).cuda()
outs = [torch.empty(S * B, H // 2, H // 2, 128, dtype=torch.half, device="cuda") for _ in range(G)]
ys = (ctypes.c_void_p * G)(*[t.data_ptr() for t in outs])
def t(fn, n=10):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
plain = lambda: lib.bnn_conv2d_tc_grouped(copies.data_ptr(), w.data_ptr(), b.data_ptr(), ys, G, 3, 0, 1, S * B, H, H, C, 128, 3, 2, stream())
masked = lambda: lib.bnn_conv2d_tc_grouped_masked(xs.data_ptr(), bits.data_ptr(), w.data_ptr(), b.data_ptr(), ys, G, 3, 1, S * B, B, H, H, C, 128, stream())
print("plain grouped on copies   %.3f ms" % t(plain))
for exp in ("0", "12", "16", "32", "48", "60"):
    os.environ["BNN_TC_EXP"] = exp
    print("masked EXP=%-2s (4 no RMW, 8 no proxy fence, 16 cta-scope wait, 32 cta-scope arrive)  %.3f ms" % (exp, t(masked)))
os.environ.pop("BNN_TC_EXP")
print("boundary_bits %.3f ms   dropout %.3f ms" % (
    t(lambda: lib.bnn_boundary_bits(x.data_ptr(), xs.data_ptr(), bits.data_ptr(), plane.data_ptr(), 1, B, H, H, C, S, ctypes.byref(dd), stream())),
    t(lambda: lib.bnn_dropout(x.data_ptr(), copies.data_ptr(), 1, H * H * C, C, S, 0, ctypes.byref(dd), stream()))))
